"""Per-layer roofline of a bench line: algorithmic FLOPs, the MMAs the tensor pipe actually executes (operand passes of
the precision, fast-FIR split of layer 8, channel / time padding of the tiles), their time at the MEASURED cuBLAS peak,
and the CUDA-event time of the launch inside the timed steps.

  python tools/layer_roofline.py profiles/r02_bench_cfg2.json [MEASURED_PEAKS.json]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAYERS = [(48, 2, 128, 250), (7, 1, 250, 250)] + [(7, 1, 250, 250)] * 6 + [(32, 1, 250, 2000), (1, 1, 2000, 2000),
                                                                            (1, 1, 2000, 29)]


def up(x, m):
  return (x + m - 1) // m * m


def main():
  line = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  peaks_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'MEASURED_PEAKS.json')
  peak = 1575.6
  if os.path.exists(peaks_path):
    pk = json.load(open(peaks_path))
    peak = float(pk.get('bf16_tflops', peak))
  r = line['roofline']
  ms = r['layers_ms_per_step']
  passes = int(r.get('mma_passes', 1))
  cfg = line['config']
  B = int(cfg['global_batch']) // max(1, int(line.get('n_gpus', 1)))
  T = int(round(float(cfg['mean_frames'])))
  rows = []
  t = T
  print('# %s' % sys.argv[1])
  print('# peak = %.1f TFLOP/s (cuBLAS bf16, measured); %d operand pass(es) per algorithmic MAC; B = %d, T = %d' % (peak, passes, B, T))
  print('%-14s %10s %12s %10s %10s %8s' % ('launch', 'alg GFLOP', 'exec GFLOP', 'ideal us', 'event us', 'ideal/ev'))
  tot = [0.0, 0.0, 0.0, 0.0]
  for l, (K, s, cin, cout) in enumerate(LAYERS):
    to = (t + s - 1) // s
    alg = 2.0 * K * cin * cout * to * B
    # executed: rows padded to 128-row tiles per utterance, output channels to 16, contraction to 16 per tap chunk
    rows_p = up(to, 128)
    cin_p = up(cin, 16) if cin % 64 else cin
    cout_p = up(cout, 16)
    ffa = 0.5625 if (l == 8 and passes <= 3) else 1.0
    for kind in ('fwd', 'dgrad', 'wgrad'):
      key = 'L%d.%s' % (l, kind)
      if key not in ms:
        continue
      m = ms[key]
      n_layers = 7 if key == 'L1.wgrad' and 'L2.wgrad' not in ms else 1       # merged launch of layers 1-7
      ex = 2.0 * K * cin_p * cout_p * (rows_p if kind != 'wgrad' else up(to, 64)) * B * passes * ffa * n_layers
      if l == 8 and ffa < 1.0:
        # quarter-rate sequences: 127 rows per utterance in ONE 128-row tile instead of 501 rows in four
        ex = 2.0 * (K // 4) * cin_p * cout_p * 128 * B * passes * 9
      ideal = ex / (peak * 1e12) * 1e6
      rows.append((key + (' (x7)' if n_layers == 7 else ''), alg * n_layers / 1e9, ex / 1e9, ideal, m * 1e3, ideal / (m * 1e3)))
      tot[0] += alg * n_layers; tot[1] += ex; tot[2] += ideal; tot[3] += m * 1e3
    t = to
  for row in rows:
    print('%-14s %10.1f %12.1f %10.1f %10.1f %8.2f' % row)
  print('%-14s %10.1f %12.1f %10.1f %10.1f %8.2f' % ('total', tot[0] / 1e9, tot[1] / 1e9, tot[2], tot[3], tot[2] / tot[3]))
  print('# event us = CUDA events around the launch in the instrumented pass (adds ~10 us of stream commands per launch'
        ' and, with the side stream, the background kernels that run beside layers 0-7); step = %.3f ms' % line['ms_per_step'])


if __name__ == '__main__':
  main()
