#!/usr/bin/env python3
"""bench.py -- utterances/s of one Wav2Letter train step (BASELINE.json metric) on N B200s of one box.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16x3|fp32|bf16] [--impl ours|reference]
  N>1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
            bench.py --gpus N --steps K --warmup W

A "step" is model.step(update=True): forward through the 11 conv layers, CTC loss, backward, [NCCL allreduce of the
flat gradient], clip_by_global_norm, Adam -- on one batch of synthetic N(0,1) 128-mel inputs.  Workload at every N is
BASELINE.json configs[1] per GPU: batch 32 x 10 s @ 16 kHz (T=1001 mel frames -> T'=501 logit frames), 11-layer net
(weak scaling: per-GPU batch fixed).

Printed by rank 0: ONE JSON line, see the keys below; `value` is the device-resident throughput, `e2e` the same step
driven through the reference-facing SpeechModel.step with HOST batches (H2D of the inputs and D2H of the loss inside
the timed region).  `--impl reference` times the CPU restatement of the reference (oracle/, numpy on all host cores;
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

METRIC = 'utterances/sec (10s@16kHz, 128-mel) train-step'
UNIT = 'utterances/s'


def parse_args():
  p = argparse.ArgumentParser()
  p.add_argument('--gpus', type=int, default=1)
  p.add_argument('--steps', type=int, default=10)
  p.add_argument('--warmup', type=int, default=3)
  p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  p.add_argument('--precision', default=os.environ.get('SPEECHT_B200_PRECISION', 'bf16x3'),
                 choices=['fp32', 'bf16x6', 'bf16x3', 'bf16'])
  p.add_argument('--batch', type=int, default=32, help='per-GPU batch (BASELINE configs[1]: 32)')
  p.add_argument('--seconds', type=float, default=10.0)
  p.add_argument('--cpu-sample', type=int, default=2, help='utterances per CPU-baseline step')
  p.add_argument('--no-cpu-baseline', action='store_true')
  return p.parse_args()


# ------------------------------------------------------------------------------------------------ workload
def frames_for(seconds):
  return 1 + int(16000 * seconds) // 160


def make_labels(rng, n_chars, ctc_len):
  while True:
    lab = rng.integers(0, 28, size=n_chars)
    if n_chars + int(np.sum(lab[1:] == lab[:-1])) <= ctc_len:
      return lab.astype(np.int32)


def make_batch(seed, batch, seconds):
  """BASELINE.md synthetic inputs: N(0,1) mel [B,T,128] f32, 15 chars/s labels feasible for CTC."""
  rng = np.random.default_rng(seed)
  T = frames_for(seconds)
  inputs = rng.standard_normal((batch, T, 128), dtype=np.float32)
  lengths = np.full((batch,), T, dtype=np.int32)
  labels = [make_labels(rng, int(15 * seconds), T // 2) for _ in range(batch)]
  return inputs, lengths, labels


def conv_flops_forward(batch, T):
  from speecht_b200.engine import layer_table
  total, per_layer, t = 0.0, [], T
  for (k, s, cin, cout, _r) in layer_table():
    t = -(-t // s)
    f = 2.0 * k * cin * cout * t * batch
    per_layer.append(f)
    total += f
  return total, per_layer


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

  def stop(self, first=0, last=None):
    self.lines = self.lines[first:last]
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    sm, mx, power, reasons = [], [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for line in self.lines:
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


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_steps(sample, seconds, steps, warmup):
  """Times oracle.train_step (float32 numpy, BLAS on every host core) on `sample` utterances per step."""
  from oracle import speecht_oracle as O
  inputs, lengths, labels = make_batch(1000, sample, seconds)
  weights = O.xavier_weights(np.random.default_rng(0), dtype=np.float32)
  m = [(np.zeros_like(w), np.zeros_like(b)) for w, b in weights]
  v = [(np.zeros_like(w), np.zeros_like(b)) for w, b in weights]
  times = []
  for i in range(warmup + steps):
    t0 = time.perf_counter()
    O.train_step(inputs, lengths, labels, weights, m, v, step=i + 1, lr=1e-4, dtype=np.float32)
    dt = time.perf_counter() - t0
    if i >= warmup:
      times.append(dt)
  sec = float(np.mean(times))
  return sample / sec, sec


def host_cores():
  try:
    return len(os.sched_getaffinity(0))
  except AttributeError:
    return os.cpu_count() or 1


def run_reference(args, rank, world):
  if rank != 0:
    return
  T = frames_for(args.seconds)
  steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
  value, sec = cpu_reference_steps(args.cpu_sample, args.seconds, steps, warmup)
  line = {
    'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
    'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
    'dtype': 'fp32', 'data': 'synthetic',
    'config': {'workload': 'configs[1] train-step: 10s@16kHz synthetic 128-mel, Wav2Letter 11 conv layers, fp32; '
                           'CPU sample of %d utterances/step (T=%d)' % (args.cpu_sample, T),
               'global_batch': args.cpu_sample, 'seconds': args.seconds},
    'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': host_cores(), 'kind': 'port',
                     'sample': '%d steps of oracle.train_step on %d x %gs utterances (numpy/BLAS float32); CPU '
                               'restatement of the reference, NOT TensorFlow-1 (not installable, DESIGN.md)'
                               % (steps, args.cpu_sample, args.seconds)},
    'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    'gpu_launches': 0,
  }
  print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, local_rank, world):
  import torch
  import torch.distributed as dist
  from speecht_b200 import speech_input, speech_model
  from speecht_b200.engine import W2LEngine

  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  group = None
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
    group = dist.group.WORLD

  B, T = args.batch, frames_for(args.seconds)
  n_sets = 4                                                 # rotate distinct batches so inputs never sit in L2
  host_sets = [make_batch(100 * rank + i, B, args.seconds) for i in range(n_sets)]
  pinned = [torch.from_numpy(h[0]).pin_memory() for h in host_sets]
  dev_inputs = [p.to(dev) for p in pinned]

  eng = W2LEngine(precision=args.precision, device=dev, process_group=group)
  eng.init_xavier(seed=0)                                    # same weights on every rank
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
    ms = e0.elapsed_time(e1)
    if world > 1:
      t = torch.tensor([ms], device=dev, dtype=torch.float64)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t.item())
    return ms

  # ---- device-resident throughput (`value`)
  def step_resident(i):
    h = host_sets[i % n_sets]
    eng.train_step(dev_inputs[i % n_sets], h[1], h[2], lr)

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
  # launch stream -- the ~70 extra stream commands per step cost a few % and are kept out of `value`
  eng.start_kernel_timing()
  ms_instrumented = timed(step_resident, args.steps) / args.steps
  timings = eng.stop_kernel_timing()

  # ---- roofline of the dominant kernel, from events recorded inside the timed region
  roofline = None
  if rank == 0:
    roofline = eng.roofline_report(timings, args.steps, os.path.join(ROOT, 'MEASURED_PEAKS.json'))
    if roofline is not None:
      roofline['instrumented_ms_per_step'] = ms_instrumented
    if roofline is not None and os.path.exists(os.path.join(ROOT, 'profiles', 'traffic.json')):
      # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
      # `ncu --set full` capture (profiles/), keyed by precision
      tr = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(args.precision, {})
      roofline['traffic'] = tr.get('bytes_per_launch')
      roofline['traffic_source'] = tr.get('source')

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
  flags = types.SimpleNamespace(command='train', learning_rate=lr, learning_rate_decay_factor=0.0,
                                max_gradient_norm=5.0, momentum=0.9, log_dir='log', run_name='bench',
                                run_type='train', precision=args.precision, process_group=group, engine=eng)
  feed = HostFeed()
  model = speech_model.create_default_model(flags, 128, feed)
  sess = speech_model.Session(dev)
  for i in range(args.warmup):
    model.step(sess)
  e2e_ms = timed(lambda i: model.step(sess), args.steps) / args.steps
  e2e_value = world * B / (e2e_ms * 1e-3)
  clocks = None
  if rank == 0:
    mark2 = sampler.mark()
    # samples taken during the device-resident timed region; if it was shorter than the 100 ms sampling period,
    # fall back to everything up to the end of the e2e region (same kernels, same load)
    clocks = sampler.stop(mark0, mark1 if mark1 > mark0 else mark2)

  aux = None
  if rank == 0 and world == 1:
    aux = aux_kernels(eng, dev, B, args.seconds, host_sets[0], os.path.join(ROOT, 'MEASURED_PEAKS.json'))
  if world > 1:
    dist.destroy_process_group()
  if rank != 0:
    return

  fwd_flops, _ = conv_flops_forward(B, T)
  line = {
    'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
    'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
    'dtype': {'fp32': 'fp32', 'bf16x3': 'fp32 (bf16x3 split operands on tcgen05, fp32 accumulate; CTC/Adam fp32)',
              'bf16x6': 'fp32 (bf16x6: three bf16 planes per operand, 6 products on tcgen05, fp32 accumulate)',
              'bf16': 'bf16 conv operands, fp32 accumulate; CTC/Adam fp32'}[args.precision],
    'data': 'synthetic',
    'config': {'workload': 'configs[1] train-step: batch %d/GPU, %gs@16kHz synthetic 128-mel (T=%d, T\'=%d), '
                           'Wav2Letter 11 conv layers' % (B, args.seconds, T, (T + 1) // 2),
               'global_batch': world * B, 'precision': args.precision, 'parallelism': 'dp%d' % world,
               'l2': 'no explicit flush: per-step working set (activations+params+grads+Adam > 1 GB) exceeds the '
                     '126 MB L2 and %d distinct input batches rotate' % n_sets,
               'train_tflop_per_step_per_gpu': 3 * fwd_flops / 1e12},
    'clocks': clocks,
    'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': e2e_ms,
            'h2d_bytes_per_step': int(pinned[0].numel() * 4 + sum(len(l) for l in host_sets[0][2]) * 4 + 8 * B + 4),
            'd2h_bytes_per_step': 4},
    'gpu_launches': int(launches),
    'roofline': roofline,
    'aux_hbm_kernels': aux,
  }
  if world == 1 and not args.no_cpu_baseline:
    cv, csec = cpu_reference_steps(args.cpu_sample, args.seconds, 2, 1)
    line['cpu_baseline'] = {'value': cv, 'unit': UNIT, 'cores': host_cores(), 'kind': 'port',
                            'sample': '2 steps of oracle.train_step on %d x %gs utterances (numpy/BLAS float32, '
                                      '%.1f s/step); CPU restatement, not TensorFlow-1' % (args.cpu_sample,
                                                                                         args.seconds, csec)}
  print(json.dumps(line), flush=True)


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
    'sumsq+clip_adam (a10-a11)': (lambda: (ops.global_norm_sq(scratch[1], nsq),
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
