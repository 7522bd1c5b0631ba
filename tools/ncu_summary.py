"""Summarises `ncu --set full` captures for profiles/: one text summary per report and profiles/traffic.json
(dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, which bench.py copies into
roofline.traffic).

  python tools/ncu_summary.py bf16x3=gpurun_out/prof_conv_v6.ncu-rep bf16=gpurun_out/prof_conv_bf16_v6.ncu-rep \
         wgrad=gpurun_out/prof_wgrad_v6.ncu-rep --round r01
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']

UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def read(rep):
  if rep.endswith('.csv'):          # raw page exported on the GPU box (tools/gpu_round2.sh)
    out = open(rep).read()
  else:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  hdr, units = rows[0], rows[1]
  res = []
  for r in rows[2:]:
    d = {}
    for h, u, v in zip(hdr, units, r):
      d[h] = (v, u)
    res.append(d)
  return res


def main():
  rnd = 'r01'
  args = [a for a in sys.argv[1:] if '=' in a]
  if '--round' in sys.argv:
    rnd = sys.argv[sys.argv.index('--round') + 1]
  traffic = {}
  for a in args:
    tag, rep = a.split('=', 1)
    launches = read(rep)
    lines = []
    for d in launches:
      lines.append('kernel: %s' % d['Kernel Name'][0][:150])
      for k in KEYS:
        if k in d:
          lines.append('  %-72s %s %s' % (k, d[k][0], d[k][1]))
      rd = float(d['dram__bytes_read.sum'][0]) * UNIT.get(d['dram__bytes_read.sum'][1], 1)
      wr = float(d['dram__bytes_write.sum'][0]) * UNIT.get(d['dram__bytes_write.sum'][1], 1)
      lines.append('  %-72s %.1f MB' % ('traffic = dram read + write', (rd + wr) / 1e6))
      lines.append('')
      traffic.setdefault(tag, {'bytes_per_launch': rd + wr, 'kernel': d['Kernel Name'][0][:80],
                               'duration_us_under_ncu': float(d['gpu__time_duration.sum'][0]),
                               'source': 'profiles/%s_ncu_%s.txt (ncu --set full, one launch)' % (rnd, tag)})
    path = os.path.join(ROOT, 'profiles', '%s_ncu_%s.txt' % (rnd, tag))
    open(path, 'w').write('\n'.join(lines))
    print('wrote', path)
  tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
  old = json.load(open(tpath)) if os.path.exists(tpath) else {}
  old.update(traffic)
  json.dump(old, open(tpath, 'w'), indent=1)
  print('wrote', tpath)


if __name__ == '__main__':
  main()
