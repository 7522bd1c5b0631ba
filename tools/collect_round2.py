"""Copies the artefacts of tools/gpu_round2.sh (gpurun_out/final/) into profiles/ (tracked): bench lines, launch list,
`ncu --set full` summaries, and profiles/traffic.json keyed by the benched shape ("<precision>/B<batch>/T<frames>") so
that bench.py only reports a measured `roofline.traffic` for a shape that has a capture.

  python tools/collect_round2.py [gpurun_out/final]
"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import ncu_summary  # noqa: E402

# capture name -> (summary tag, shape key or None, which launch it is)
CAPTURES = {
  'l8_fwd': ('l8_fwd_bf16x3', 'bf16x3/B32/T1001',
             'layer-8 forward: the nine-problem fast-FIR launch on CTA pairs, the longest tc_conv_kernel launch of the step'),
  'l8_dgrad': ('l8_dgrad_bf16x3', None, 'layer-8 data gradient (nine-problem launch on CTA pairs)'),
  'l9_dgrad': ('l9_dgrad_bf16x3', None, 'layer-9 data gradient on CTA pairs'),
  'l1_fwd': ('l1_fwd_bf16x3', None, 'layer-1 forward (250 channels, one tile per CTA)'),
  'l8_wgrad': ('l8_wgrad_bf16x3', None, 'layer-8 filter gradient (nine-problem launch)'),
  'l1_7_wgrad': ('l1_7_wgrad_bf16x3', None, 'filter gradients of the seven 250-channel layers, one launch, forced K slices'),
  'ctc_alpha_beta': ('ctc_alpha_beta', None, 'CTC alpha/beta recursion'),
  'pack': ('pack_bwd', None, 'streaming pack of the nine leaf filters of layer 8 (backward layout only; the forward kernel reads it MN-major)'),
  'l10_dgrad': ('l10_dgrad_bf16x3', None, 'layer-10 data gradient: 128-wide tiles, sixteen epilogue warps'),
  'ffa2_combine': ('ffa2_combine', None, 'forward combine of the nine leaf products (bias + ReLU + plane split)'),
  'ffa2_dz_prep': ('ffa2_dz_prep', None, 'backward prepare: the nine leaf gradients from dy'),
  'l8_fwd_cfg3': ('l8_fwd_bf16_cfg3', 'bf16/B64/T1001', 'layer-8 forward at config 3 (plain bf16, batch 64)'),
}


def main():
  src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'final')
  prof = os.path.join(ROOT, 'profiles')
  for name in sorted(os.listdir(src)):
    path = os.path.join(src, name)
    if name.startswith('bench_') and name.endswith('.json') and os.path.getsize(path) > 0:
      shutil.copy(path, os.path.join(prof, 'r02_' + name))
    elif name == 'launches_cfg2.csv':
      shutil.copy(path, os.path.join(prof, 'r02_launches_cfg2.csv'))
    elif name == 'summary.txt':
      shutil.copy(path, os.path.join(prof, 'r02_final_session.txt'))
    elif name == 'accuracy_sweep.txt':
      shutil.copy(path, os.path.join(prof, 'r02_accuracy_sweep_8seeds_final.txt'))
    elif name == 't_all.log':
      keep = [l for l in open(path) if ('passed' in l or 'vs float64' in l or 'near-tie' in l or 'fast-FIR level' in l)]
      open(os.path.join(prof, 'r02_gpu_tests_final.txt'), 'w').write(''.join(keep))
  traffic = {}
  for cap, (tag, key, what) in CAPTURES.items():
    rep = os.path.join(src, 'ncu_%s.raw.csv' % cap)
    if not os.path.exists(rep):
      rep = os.path.join(src, 'ncu_%s.ncu-rep' % cap)
    if not os.path.exists(rep):
      continue
    launches = ncu_summary.read(rep)
    lines = ['# %s' % what]
    for d in launches:
      lines.append('kernel: %s' % d['Kernel Name'][0][:150])
      for k in ncu_summary.KEYS:
        if k in d:
          lines.append('  %-72s %s %s' % (k, d[k][0], d[k][1]))
      rd = float(d['dram__bytes_read.sum'][0]) * ncu_summary.UNIT.get(d['dram__bytes_read.sum'][1], 1)
      wr = float(d['dram__bytes_write.sum'][0]) * ncu_summary.UNIT.get(d['dram__bytes_write.sum'][1], 1)
      lines.append('  %-72s %.1f MB' % ('traffic = dram read + write', (rd + wr) / 1e6))
      if key:
        traffic[key] = {'bytes_per_launch': rd + wr, 'kernel': d['Kernel Name'][0][:80], 'launch': what,
                        'duration_us_under_ncu': float(d['gpu__time_duration.sum'][0]),
                        'source': 'profiles/r02_ncu_%s.txt (ncu --set full, one launch)' % tag}
    open(os.path.join(prof, 'r02_ncu_%s.txt' % tag), 'w').write('\n'.join(lines) + '\n')
    print('wrote r02_ncu_%s.txt' % tag)
  tpath = os.path.join(prof, 'traffic.json')
  json.dump(traffic, open(tpath, 'w'), indent=1)
  print('wrote', tpath, sorted(traffic))


if __name__ == '__main__':
  main()
