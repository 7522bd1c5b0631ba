#!/usr/bin/env python3
"""bench.py -- the reference's headline workloads (BASELINE.json `configs`) on N B200s of one box.

  python bench.py [--config 2|3|4|5] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  N>1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
            bench.py --gpus N --steps K --warmup W [--config C]

  --config 2 (default)  train step, batch 32/GPU x 10 s, 11-layer Wav2Letter, fp32-grade arithmetic (bf16x3 split)
  --config 3            train step, batch 64/GPU x 10 s, bf16 operands + fp32 accumulate / CTC / Adam
                        (the "17-layer large" net of BASELINE.json does not exist in the reference: SURVEY.md 0.3 #2;
                        the reference's only network, 11 layers with 2000-channel tail, is what runs)
  --config 4            train step, batch 32/GPU x 30 s, bf16, data parallel (NCCL gradient allreduce)
  --config 5            evaluate step (forward + CTC loss + greedy decode), batch 256/GPU, lengths uniform in 1..30 s
  --batch / --seconds / --precision override the config's values.

A train "step" is model.step(update=True): forward through the 11 conv layers, CTC loss, backward, [NCCL allreduce of
the flat gradient], clip_by_global_norm, Adam.  An evaluate "step" is model.step(update=False, decode=True,
return_label=True) (evaluation.py:132-137 in the reference): ONE forward, the CTC loss and the greedy decode.

Rank 0 prints ONE JSON line.  `value` is device-resident throughput (inputs already in HBM), `e2e` the same step through
the reference-facing SpeechModel.step with HOST batches (H2D of the inputs and D2H of the result inside the timed
region).  `ctc_loss_delta` compares the GPU path with the float64 oracle on a small sample of the same workload.
`roofline` is measured live with CUDA events around the dominant kernel's launches; its `frac` uses the burst cuBLAS
figure when the timed region is shorter than a second and the sustained one otherwise (both fractions are reported).
`--impl reference` times the CPU restatement of the reference (oracle/, numpy on a pinned number of host threads;
TensorFlow 1.x cannot be installed here) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = 'utterances/s'
PEAKS_PATH = os.path.join(ROOT, 'MEASURED_PEAKS.json')
TRAFFIC_PATH = os.path.join(ROOT, 'profiles', 'traffic.json')

CONFIGS = {
  2: dict(mode='train', batch=32, seconds=10.0, precision='bf16x3',
          name='configs[1] train-step: batch 32/GPU, 10s@16kHz synthetic 128-mel, Wav2Letter 11 conv layers, fp32-grade '
               '(bf16x3 split operands)'),
  3: dict(mode='train', batch=64, seconds=10.0, precision='bf16',
          name='configs[2] train-step: batch 64/GPU, 10s@16kHz, bf16 conv stack + fp32 CTC (the reference has ONE '
               'network: 11 conv layers, 2000-channel tail; a 17-layer net does not exist in it)'),
  4: dict(mode='train', batch=32, seconds=30.0, precision='bf16',
          name='configs[3] data-parallel train-step: batch 32/GPU, 30s utterances, bf16, NCCL gradient allreduce'),
  5: dict(mode='eval', batch=256, seconds=(1, 30), precision='bf16x3',
          name='configs[4] evaluate: batch 256/GPU, greedy CTC decode, variable-length 1-30s synthetic utterances '
               '(zero-padded to the batch maximum like speech_input.py:38-43; unbucketed)'),
}


def parse_args():
  p = argparse.ArgumentParser()
  p.add_argument('--gpus', type=int, default=1)
  p.add_argument('--steps', type=int, default=10)
  p.add_argument('--warmup', type=int, default=3)
  p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  p.add_argument('--config', type=int, default=2, choices=sorted(CONFIGS))
  p.add_argument('--precision', default=os.environ.get('SPEECHT_B200_PRECISION'),
                 choices=['fp32', 'bf16x6', 'bf16x3', 'bf16'])
  p.add_argument('--batch', type=int, default=None, help='per-GPU batch (default: the config\'s)')
  p.add_argument('--seconds', type=float, default=None, help='fixed utterance length (default: the config\'s)')
  p.add_argument('--cpu-sample', type=int, default=8, help='utterances per CPU-baseline step')
  p.add_argument('--cpu-threads', type=int, default=16, help='BLAS threads of the CPU arm (capped by the host)')
  p.add_argument('--no-cpu-baseline', action='store_true')
  p.add_argument('--no-sustained', action='store_true', help='skip the >= 2 s sustained pass')
  p.add_argument('--sustained-seconds', type=float, default=2.5)
  p.add_argument('--eval-buckets', type=int, default=1,
                 help='config 5 only: evaluate each batch as this many length-sorted groups padded to their own maximum '
                      '(default 1 = the reference behaviour: the whole batch padded to its longest utterance)')
  args = p.parse_args()
  cfg = dict(CONFIGS[args.config])
  if args.batch is not None:
    cfg['batch'] = args.batch
  if args.seconds is not None:
    cfg['seconds'] = args.seconds
  if args.precision is not None:
    cfg['precision'] = args.precision
  args.cfg = cfg
  return args


# ------------------------------------------------------------------------------------------------ workload
def frames_for(seconds):
  return 1 + int(16000 * seconds) // 160


def make_labels(rng, n_chars, ctc_len):
  while True:
    lab = rng.integers(0, 28, size=n_chars)
    if n_chars + int(np.sum(lab[1:] == lab[:-1])) <= ctc_len:
      return lab.astype(np.int32)


def make_batch(seed, batch, seconds):
  """BASELINE.md synthetic inputs: N(0,1) mel [B,T,128] f32, 15 chars/s labels feasible for CTC.
  seconds: a number (fixed length) or (lo, hi): integer seconds uniform in [lo, hi], zero-padded to the batch max."""
  rng = np.random.default_rng(seed)
  if isinstance(seconds, (tuple, list)):
    secs = rng.integers(int(seconds[0]), int(seconds[1]) + 1, size=batch).astype(np.float64)
  else:
    secs = np.full((batch,), float(seconds))
  lengths = np.array([frames_for(s) for s in secs], dtype=np.int32)
  T = int(lengths.max())
  inputs = np.zeros((batch, T, 128), dtype=np.float32)
  for b in range(batch):
    inputs[b, :lengths[b]] = rng.standard_normal((int(lengths[b]), 128), dtype=np.float32)
  labels = [make_labels(rng, int(15 * secs[b]), int(lengths[b]) // 2) for b in range(batch)]
  return inputs, lengths, labels


def conv_flops_forward(batch, T):
  from speecht_b200.engine import layer_table
  total, t = 0.0, T
  for (k, s, cin, cout, _r) in layer_table():
    t = -(-t // s)
    total += 2.0 * k * cin * cout * t * batch
  return total


def n_input_sets(cfg):
  return 4 if cfg['mode'] == 'train' else 2              # distinct batches rotated so that inputs never sit in L2


def batch_lengths(seed, batch, seconds):
  """The frame counts make_batch(seed, batch, seconds) draws (its first use of the generator)."""
  rng = np.random.default_rng(seed)
  if isinstance(seconds, (tuple, list)):
    secs = rng.integers(int(seconds[0]), int(seconds[1]) + 1, size=batch).astype(np.float64)
  else:
    secs = np.full((batch,), float(seconds))
  return np.array([frames_for(s) for s in secs], dtype=np.int32)


def workload_config(args, world):
  """`config` of the JSON line: a function of the command line only, so that the GPU arm and the reference arm
  (which times a bounded SAMPLE of the same workload, described in its cpu_baseline.sample) print the same dict."""
  cfg = args.cfg
  B = cfg['batch']
  lens = [batch_lengths(100 * 0 + i, B, cfg['seconds']) for i in range(n_input_sets(cfg))]   # rank 0's batches
  T = int(max(l.max() for l in lens))
  return {'workload': cfg['name'] + ' (T=%d mel frames -> T\'=%d logit frames)' % (T, (T + 1) // 2),
          'bench_config': args.config, 'mode': cfg['mode'], 'global_batch': world * B, 'precision': cfg['precision'],
          'parallelism': 'dp%d' % world, 'mean_frames': float(np.mean([l.mean() for l in lens])),
          'eval_buckets': args.eval_buckets,
          'l2': 'no explicit flush: per-step working set (activations+params+grads+Adam > 1 GB) exceeds the '
                '126 MB L2 and %d distinct input batches rotate' % n_input_sets(cfg),
          'conv_tflop_per_step_per_gpu': (3 if cfg['mode'] == 'train' else 1) * conv_flops_forward(B, T) / 1e12}


def metric_name(cfg):
  if cfg['mode'] == 'eval':
    return 'utterances/sec (1-30s@16kHz, 128-mel) evaluate-step: forward + CTC loss + greedy decode'
  secs = cfg['seconds']
  return 'utterances/sec (%gs@16kHz, 128-mel) train-step' % secs


def reference_weights():
  """Xavier weights from the SAME stream on the GPU arm, the CPU arm and the loss-delta check."""
  from oracle import speecht_oracle as O
  return O.xavier_weights(np.random.default_rng(0), dtype=np.float32)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
  """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
  QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
           'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
           'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index=0):
    self.gpu_index, self.proc, self.lines = gpu_index, None, []

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                                    '-lms', '25', '-i', str(self.gpu_index)], stdout=subprocess.PIPE,
                                   stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except (OSError, ValueError):
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def mark(self):
    return len(self.lines)

  def summary(self, first=0, last=None):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    sm, mx, power, reasons = [], [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for line in self.lines[first:last]:
      f = [x.strip() for x in line.split(',')]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
      except ValueError:
        continue
      for name, val in zip(names, f[5:9]):
        if val.lower().startswith('active'):
          reasons.add(name)
    if not sm:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
    return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'power_w_max': float(max(power)),
            'samples': len(sm), 'reasons': sorted(reasons)}

  def stop(self):
    if self.proc is None:
      return
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except subprocess.TimeoutExpired:
      self.proc.kill()


# ------------------------------------------------------------------------------------------------ CPU arm
def host_cores():
  try:
    return len(os.sched_getaffinity(0))
  except AttributeError:
    return os.cpu_count() or 1


def cpu_threads(requested):
  return max(1, min(int(requested), host_cores()))


def cpu_reference_steps(mode, sample, seconds, steps, warmup, threads):
  """Times the oracle's step (float32 numpy, BLAS pinned to `threads` threads) on `sample` utterances per step.
  -> (utterances/s, seconds per step)."""
  from threadpoolctl import threadpool_limits
  from oracle import speecht_oracle as O
  inputs, lengths, labels = make_batch(1000, sample, seconds)
  weights = reference_weights()
  m = [(np.zeros_like(w), np.zeros_like(b)) for w, b in weights]
  v = [(np.zeros_like(w), np.zeros_like(b)) for w, b in weights]
  times = []
  with threadpool_limits(limits=threads):
    for i in range(warmup + steps):
      t0 = time.perf_counter()
      if mode == 'eval':
        O.evaluate_step(inputs, lengths, labels, weights, dtype=np.float32)
      else:
        O.train_step(inputs, lengths, labels, weights, m, v, step=i + 1, lr=1e-4, dtype=np.float32)
      dt = time.perf_counter() - t0
      if i >= warmup:
        times.append(dt)
  sec = float(np.mean(times))
  return sample / sec, sec


def cpu_sample_seconds(cfg):
  """The CPU arm's bounded sample of the workload: fixed-length utterances; for the variable-length evaluate
  workload its mean length (15.5 s), so that utterances/s stay comparable."""
  secs = cfg['seconds']
  return float(np.mean(secs)) if isinstance(secs, (tuple, list)) else float(secs)


def cpu_baseline_record(cfg, sample, steps, warmup, threads):
  secs = cpu_sample_seconds(cfg)
  value, sec = cpu_reference_steps(cfg['mode'], sample, secs, steps, warmup, threads)
  what = 'oracle.evaluate_step' if cfg['mode'] == 'eval' else 'oracle.train_step'
  return {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'host_cores': host_cores(),
          'sample': '%d steps of %s on %d x %gs utterances (numpy/OpenBLAS float32 pinned to %d threads, %.2f s/step); '
                    'CPU restatement of the reference, NOT TensorFlow-1 (not installable, DESIGN.md)'
                    % (steps, what, sample, secs, threads, sec)}, sec


def run_reference(args, rank, world):
  if rank != 0:
    return
  cfg = args.cfg
  threads = cpu_threads(args.cpu_threads)
  secs = cpu_sample_seconds(cfg)
  # bounded: about 1.2 s per train step at 8 x 10 s on 16 threads (slower boxes: 3 s); K and W are honoured as long as
  # the whole run stays near two minutes
  per_step_guess = 1.5 * (args.cpu_sample / 8.0) * (secs / 10.0) * (0.4 if cfg['mode'] == 'eval' else 1.0)
  warmup = int(max(0, min(args.warmup, 5)))
  steps = int(max(1, min(args.steps, 120.0 / max(per_step_guess, 1e-3) - warmup)))
  rec, sec = cpu_baseline_record(cfg, args.cpu_sample, steps, warmup, threads)
  line = {
    'impl': 'reference', 'metric': metric_name(cfg), 'value': rec['value'], 'unit': UNIT, 'n_gpus': args.gpus,
    'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
    'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
    'config': workload_config(args, max(1, args.gpus)),
    'cpu_baseline': rec,
    'e2e': {'value': rec['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    'gpu_launches': 0,
  }
  print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ parity beside the number
def ctc_loss_delta(eng_factory, cfg, sample=2):
  """|CTC loss(GPU) - CTC loss(float64 oracle)| on `sample` utterances of the CPU arm's own batch, same weights
  (BASELINE.json metric: "CTC-loss delta vs ref").  Greedy labels are compared too (bit-exact gate)."""
  import torch
  from oracle import speecht_oracle as O
  secs = min(cpu_sample_seconds(cfg), 10.0)          # bounded: the float64 oracle forward is ~1 s per 10 s utterance
  inputs, lengths, labels = make_batch(1000, sample, secs)
  weights = reference_weights()
  t0 = time.perf_counter()
  ref = O.evaluate_step(inputs, lengths, labels, weights, dtype=np.float64)
  t_ref = time.perf_counter() - t0
  eng = eng_factory()
  eng.load_weights(weights)
  res = eng.evaluate_step(torch.from_numpy(inputs).to(eng.device), lengths, labels, decode=True)
  loss = res['loss'].cpu().numpy().astype(np.float64)
  logits = res['logits'].cpu().numpy().astype(np.float64)
  d = np.abs(loss - ref['loss'])
  out = {'max_abs': float(d.max()), 'max_rel': float((d / np.abs(ref['loss'])).max()),
         'avg_loss_gpu': float(loss.mean()), 'avg_loss_oracle': float(ref['loss'].mean()),
         'logits_max_rel': float(np.max(np.abs(logits - ref['logits'])) / np.max(np.abs(ref['logits']))),
         'greedy_labels_equal': bool(np.array_equal(res['decoded'][0].values, ref['decoded'][1])
                                     and np.array_equal(res['decoded'][0].indices, ref['decoded'][0])),
         'sample': '%d x %gs utterances, float64 numpy oracle (%.1f s), precision %s' % (sample, secs, t_ref,
                                                                                        eng.precision),
         'gate': '1e-4 relative (north_star); greedy labels bit-exact'}
  del eng
  torch.cuda.empty_cache()
  return out


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, local_rank, world):
  import torch
  import torch.distributed as dist
  from speecht_b200 import parallel, speech_input, speech_model
  from speecht_b200.engine import W2LEngine

  cfg = args.cfg
  mode, precision = cfg['mode'], cfg['precision']
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  group = None
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
    group = dist.group.WORLD

  B = cfg['batch']
  n_sets = n_input_sets(cfg)
  host_sets = [make_batch(100 * rank + i, B, cfg['seconds']) for i in range(n_sets)]
  pinned = [torch.from_numpy(h[0]).pin_memory() for h in host_sets]
  dev_inputs = [p.to(dev) for p in pinned]
  T = int(max(h[0].shape[1] for h in host_sets))

  delta = None
  if rank == 0:
    delta = ctc_loss_delta(lambda: W2LEngine(precision=precision, device=dev), cfg)

  eng = W2LEngine(precision=precision, device=dev, process_group=group)
  eng.load_weights(reference_weights())                      # same weights on every rank (and on the CPU arm)
  lr = 1e-4

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(dev)

  def timed(fn, steps):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
      fn(i)
    e1.record()
    barrier()
    return parallel.max_scalar(e0.elapsed_time(e1), device=dev)

  # ---- device-resident step (`value`)
  if mode == 'train':
    def step_resident(i):
      h = host_sets[i % n_sets]
      eng.train_step(dev_inputs[i % n_sets], h[1], h[2], lr)
  else:
    from speecht_b200 import ops
    # one (inputs, CTC batch) pair per length bucket of every set (a single bucket = the reference's padded batch)
    groups = []
    for h, x in zip(host_sets, dev_inputs):
      per_set = []
      for g in W2LEngine.length_buckets(h[1], args.eval_buckets):
        t_max = max(int(h[1][g].max()), 2)
        xg = x if args.eval_buckets <= 1 else x[torch.from_numpy(g.astype(np.int64)).to(dev)][:, :t_max].contiguous()
        per_set.append((xg, ops.CTCBatch([h[2][b] for b in g], h[1][g] // 2, (xg.shape[1] + 1) // 2, eng.num_classes,
                                         dev)))
      groups.append(per_set)

    def step_resident(i):
      # forward + CTC loss + greedy-decode kernels; the compacted label rows stay on the device (the host-side
      # SparseTensor assembly and, at N > 1, the gather of the rows belong to the e2e number below)
      for xg, batch in groups[i % n_sets]:
        eng.evaluate_step_device(xg, batch)

  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()                                          # nvidia-smi needs ~1 s before its first sample
  for i in range(args.warmup):
    step_resident(i)
  # pass 1 (the headline `value`): K steps, nothing but the step's own kernels on the stream
  launches0 = eng.launches
  mark0 = sampler.mark()
  ms_total = timed(step_resident, args.steps)
  mark1 = sampler.mark()
  launches = eng.launches - launches0
  ms_step = ms_total / args.steps
  value = world * B / (ms_step * 1e-3)
  # pass 2 (roofline): the same K steps again with a CUDA-event pair recorded around every conv launch on the
  # launch stream -- the extra stream commands cost a few % and are kept out of `value`
  eng.start_kernel_timing()
  ms_instrumented = timed(step_resident, args.steps) / args.steps
  timings = eng.stop_kernel_timing()
  # pass 3 (sustained): the same step for >= 2.5 s, so the power cap has settled (MEASURED_PEAKS' sustained figure
  # is the denominator that goes with THIS number)
  sustained = None
  if not args.no_sustained:
    n_sus = int(max(args.steps, np.ceil(args.sustained_seconds * 1e3 / ms_step)))
    mark_s0 = sampler.mark()
    ms_sus = timed(step_resident, n_sus) / n_sus
    mark_s1 = sampler.mark()
    sustained = {'steps': n_sus, 'ms_per_step': ms_sus, 'value': world * B / (ms_sus * 1e-3),
                 'seconds': ms_sus * n_sus * 1e-3}
    if rank == 0:
      sustained['clocks'] = sampler.summary(mark_s0, mark_s1)

  dp_identical = None
  if world > 1 and mode == 'train':
    dp_identical = bool(parallel.identical_across_ranks(eng.params, group))

  # ---- roofline of the dominant kernel, from events recorded inside the timed region
  roofline = None
  if rank == 0:
    roofline = eng.roofline_report(timings, args.steps, PEAKS_PATH, region_seconds=ms_total * 1e-3,
                                   train=(mode == 'train'))
    if roofline is not None:
      roofline['instrumented_ms_per_step'] = ms_instrumented
      if sustained is not None:
        roofline['frac_sustained_run'] = (roofline['achieved'] * ms_step / sustained['ms_per_step']) / \
          roofline['peak_sustained'] if roofline.get('peak_sustained') else None
      roofline['traffic'], roofline['traffic_source'] = lookup_traffic(precision, B, T, roofline.get('kernel'))

  # ---- end-to-end through SpeechModel.step with host batches (`e2e`)
  class HostFeed(speech_input.BaseInputLoader):
    def __init__(self):
      super().__init__(128)
      self.inputs, self.sequence_lengths, self.labels = (speech_input.Placeholder(n) for n in
                                                         ('inputs', 'sequence_lengths', 'labels'))
      self.i = 0
      self.prefetchable = True

    def get_inputs(self):
      return self.inputs, self.sequence_lengths, self.labels

    def dequeue(self):
      k = self.i % n_sets
      self.i += 1
      return pinned[k], host_sets[k][1], host_sets[k][2]

  import types
  flags = types.SimpleNamespace(command='train' if mode == 'train' else 'evaluate', learning_rate=lr,
                                learning_rate_decay_factor=0.0, max_gradient_norm=5.0, momentum=0.9, log_dir='log',
                                run_name='bench', run_type='train', precision=precision, process_group=group,
                                engine=eng, language_model=None, eval_buckets=args.eval_buckets)
  feed = HostFeed()
  model = speech_model.create_default_model(flags, 128, feed)
  sess = speech_model.Session(dev)
  d2h_bytes = [4 * B]                                       # the [B] per-utterance losses, averaged on the host
  if mode == 'train':
    def step_e2e(i):
      model.step(sess)
  else:
    def step_e2e(i):
      _loss, decoded, _labels = model.step(sess, update=False, decode=True, return_label=True)
      rows = decoded[0]
      d2h_bytes[0] = 4 * B + 4 * B + 4 * int(rows.values.shape[0])
      if world > 1:
        parallel.gather_decoded_sparse(rows, group, device=dev)
  for i in range(args.warmup):
    step_e2e(i)
  e2e_ms = timed(step_e2e, args.steps) / args.steps
  # PCIe health beside the e2e number: pinned host -> device bandwidth of this box at this moment.  An e2e step needs
  # h2d_bytes_per_step / ms_per_step (3.8 GB/s at config 2); two runs in 30 saw ~2 GB/s on a shared host (other tenants
  # on the PCIe switch) and an e2e value bound by the upload, with `value` untouched (profiles/README.md).
  h2d_probe = None
  try:
    probe_src, probe_dst = pinned[0].view(-1), torch.empty_like(pinned[0].view(-1), device=dev)
    probe_dst.copy_(probe_src, non_blocking=True)
    torch.cuda.synchronize(dev)
    t_probe = time.perf_counter()
    for _ in range(3):
      probe_dst.copy_(probe_src, non_blocking=True)
    torch.cuda.synchronize(dev)
    h2d_probe = 3 * probe_src.numel() * 4 / 1e9 / (time.perf_counter() - t_probe)
    del probe_dst
  except Exception:
    h2d_probe = None
  e2e_value = world * B / (e2e_ms * 1e-3)
  clocks = None
  if rank == 0:
    mark2 = sampler.mark()
    # samples taken during the device-resident timed region; if it was shorter than the sampling period, everything
    # up to the end of the e2e region (same kernels, same load)
    clocks = sampler.summary(mark0, mark1 if mark1 > mark0 else mark2)
    sampler.stop()

  aux = None
  if rank == 0 and world == 1 and mode == 'train':
    aux = aux_kernels(eng, dev, B, cfg['seconds'], host_sets[0], PEAKS_PATH)
  if world > 1:
    dist.destroy_process_group()
  if rank != 0:
    return

  # inputs: SpeechModel uploads only the valid frames of a batch that is more than a quarter padding (the zeros are
  # written on the device), the whole padded tensor otherwise
  full_bytes, valid_bytes = int(pinned[0].numel() * 4), int(host_sets[0][1].sum()) * 128 * 4
  h2d_inputs = valid_bytes if valid_bytes < 0.75 * full_bytes else full_bytes
  h2d = int(h2d_inputs + sum(len(l) for l in host_sets[0][2]) * 4 + 8 * B + 4)
  line = {
    'metric': metric_name(cfg), 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
    'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
    'dtype': {'fp32': 'fp32', 'bf16x3': 'fp32 (bf16x3 split operands on tcgen05, fp32 accumulate; CTC/Adam fp32)',
              'bf16x6': 'fp32 (bf16x6: three bf16 planes per operand, 6 products on tcgen05, fp32 accumulate)',
              'bf16': 'bf16 conv operands, fp32 accumulate; CTC/Adam fp32'}[precision],
    'data': 'synthetic',
    'config': workload_config(args, world),
    'clocks': clocks,
    'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': e2e_ms, 'h2d_bytes_per_step': h2d,
            'd2h_bytes_per_step': int(d2h_bytes[0]), 'h2d_probe_GB/s': h2d_probe},
    'gpu_launches': int(launches),
    'ctc_loss_delta': delta,
    'roofline': roofline,
    'sustained': sustained,
    'aux_hbm_kernels': aux,
  }
  if dp_identical is not None:
    line['dp_identical'] = dp_identical
  if world == 1 and not args.no_cpu_baseline:
    rec, _sec = cpu_baseline_record(cfg, args.cpu_sample, 2, 1, cpu_threads(args.cpu_threads))
    line['cpu_baseline'] = rec
  print(json.dumps(line), flush=True)


def lookup_traffic(precision, B, T, kernel):
  """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
  `ncu --set full` capture of THIS shape (profiles/traffic.json, key "<precision>/B<batch>/T<frames>"); None when no
  capture of this shape exists -- a number measured on another shape would be silently wrong."""
  if not os.path.exists(TRAFFIC_PATH):
    return None, None
  tr = json.load(open(TRAFFIC_PATH)).get('%s/B%d/T%d' % (precision, B, T))
  if not tr or (kernel and kernel not in tr.get('kernel', '')):
    return None, None
  return tr.get('bytes_per_launch'), '%s; launch: %s' % (tr.get('source'), tr.get('launch', 'n/a'))


def aux_kernels(eng, dev, B, seconds, host_set, peaks_path):
  """The HBM-bound rows of SURVEY.md 8(d) -- features, CTC, greedy decode, clip+Adam -- timed alone with CUDA
  events (20 calls after 3 warm-ups): algorithmic bytes / time against the measured HBM copy bandwidth.  They are
  latency/launch-bound at these sizes; the fractions are reported as measured."""
  import torch
  from speecht_b200 import ops
  peak = 6650.0
  if os.path.exists(peaks_path):
    peak = float(json.load(open(peaks_path)).get('hbm_gbs', peak))
  T = frames_for(seconds)
  To = (T + 1) // 2
  rng = np.random.default_rng(7)
  n_samp = int(16000 * seconds)
  wav = torch.from_numpy((0.1 * rng.standard_normal((B, n_samp))).astype(np.float32)).to(dev)
  logits_bm = torch.randn((B, To, 32), device=dev)[:, :, :29]
  logits = logits_bm.transpose(0, 1)
  ctc_len = host_set[1] // 2
  batch = ops.CTCBatch(host_set[2], ctc_len, To, 29, dev)
  n = eng.params.numel()

  def timeit(fn, reps=20):
    for _ in range(3):
      fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / reps

  scratch = [torch.zeros_like(eng.params) for _ in range(4)]
  nsq = torch.ones((1,), dtype=torch.float64, device=dev)
  cases = {
    'melspec (a1-a3)': (lambda: ops.power_spectrogram(wav, [n_samp] * B, 16000), B * (4 * n_samp + 4 * T * 128)),
    'ctc_loss+grad (a8-a9)': (lambda: ops.ctc_loss(batch, logits, want_grad=True), 2 * To * B * 29 * 4),
    'ctc_greedy_decode (a12)': (lambda: lib_decode(ops, logits, batch), To * B * 29 * 4),
    'sumsq+clip_adam (a10-a11)': (lambda: (ops.global_norm_sq(scratch[1], nsq, zero=True),
                                           ops.clip_adam(scratch[0], scratch[1], scratch[2], scratch[3], 1, 1e-4,
                                                         normsq=nsq)), 8 * n * 4),
  }
  out = {}
  for name, (fn, nbytes) in cases.items():
    ms = timeit(fn)
    gbs = nbytes / (ms * 1e-3) / 1e9
    out[name] = {'ms': round(ms, 4), 'algorithmic_MB': round(nbytes / 1e6, 2), 'GB/s': round(gbs, 1),
                 'frac_of_measured_hbm': round(gbs / peak, 4)}
  out['peak_GB/s'] = peak
  return out


def lib_decode(ops, logits, batch):
  """Greedy decode kernel only (no host assembly of the sparse triple)."""
  import torch
  from speecht_b200._lib import check, lib, ptr, stream_ptr
  T, B, C = logits.shape
  key = ('dec', T, B)
  buf = lib_decode.cache.get(key)
  if buf is None:
    buf = (torch.empty((B, T), dtype=torch.int32, device=logits.device),
           torch.empty((B,), dtype=torch.int32, device=logits.device),
           torch.empty((B,), dtype=torch.float32, device=logits.device))
    lib_decode.cache[key] = buf
  check(lib().st_ctc_greedy_decode(ptr(logits), logits.stride(0), logits.stride(1), T, B, C, ptr(batch.seq_len), C - 1,
                                   1, ptr(buf[0]), ptr(buf[1]), ptr(buf[2]), stream_ptr()))


lib_decode.cache = {}


def main():
  args = parse_args()
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if args.impl == 'reference':
    run_reference(args, rank, world)
    return
  if world != args.gpus:
    if world == 1 and args.gpus > 1:
      raise SystemExit('--gpus %d needs torchrun (see the module docstring)' % args.gpus)
    args.gpus = world
  run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
  main()
