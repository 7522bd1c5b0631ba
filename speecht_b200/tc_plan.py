"""TCPlan -- host side of the tensor-core step plan (csrc/w2l_plan.cu) for the 'bf16x3' and 'bf16' precisions.

One native plan per (batch, time) shape, each bound to an arena allocated through torch and to the engine's flat
parameter / gradient buffers.  The sequence per train step is
  pack_weights (fp32 -> bf16 planes)  ->  plan.forward  ->  st_ctc_loss (writes d loss/d logits planes)
  ->  plan.backward  ->  [NCCL allreduce]  ->  clip_by_global_norm + Adam on the flat fp32 buffers.
"""
import ctypes
import os

import numpy as np
import torch

from . import ops
from ._lib import check, lib, ptr, stream_ptr

N_PLANES = {'bf16': 1, 'bf16x3': 2, 'bf16x6': 3}
# SMs the backward pass of layers 7..0 leaves to the concurrent NCCL allreduce in data-parallel runs.  Measured on two
# GPUs (profiles/r02_dp_reserve_ab_2gpu.txt): NCCL finds its SMs anyway (the 250-channel layers occupy 128 of 148), and
# reserving 8 / 16 costs 1.5 % per step -- so the default is 0; the switch stays for boxes where NCCL needs more CTAs.
DP_RESERVE_SMS = int(os.environ.get('SPEECHT_B200_DP_RESERVE_SMS', '0'))
MAX_CACHED_SHAPES = 64


class _Shape:
  """Native plan of one (batch, time) shape, bound into the TCPlan's shared arena."""

  def __init__(self, engine, B, T):
    self.B, self.T = B, T
    self.handle = ctypes.c_void_p()
    check(lib().st_plan_create(ctypes.byref(self.handle), B, T, engine.input_size, engine.num_classes,
                               N_PLANES[engine.precision]))
    n = lib().st_plan_param_floats(self.handle)
    if n != engine.params.numel():
      raise RuntimeError('native parameter layout (%d floats) != engine layout (%d)' % (n, engine.params.numel()))
    self.nbytes = lib().st_plan_arena_bytes(self.handle)

  def bind(self, engine, arena):
    self.arena = arena
    nbytes = self.nbytes
    B = self.B
    base = self.arena.data_ptr()
    self.arena_ptr = (base + 1023) // 1024 * 1024
    check(lib().st_plan_bind(self.handle, ctypes.c_void_p(self.arena_ptr), nbytes, ptr(engine.params),
                             ptr(engine.grads)))
    self.To = lib().st_plan_logit_frames(self.handle)
    # fp32 logits [B, To, 32] inside the arena, exposed as the reference's time-major [T', B, C] view
    off = lib().st_plan_logits(self.handle) - base
    flat = self.arena[off:off + B * self.To * 32 * 4].view(torch.float32).view(B, self.To, 32)
    self.logits_bm = flat[:, :, :engine.num_classes]
    self.logits_tm = self.logits_bm.transpose(0, 1)
    off = lib().st_plan_dlogits_planes(self.handle) - base
    npl = N_PLANES[engine.precision]
    self.dlogits_planes = self.arena[off:off + npl * B * self.To * 64 * 2].view(torch.bfloat16).view(npl, B, self.To, 64)
    self.filter_set = lib().st_plan_filter_set(self.handle)

  def close(self):
    if self.handle:
      lib().st_plan_destroy(self.handle)
      self.handle = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass


class TCPlan:

  def __init__(self, engine):
    self.engine = engine
    self.shapes = {}
    self.arena = None                 # ONE arena shared by every shape (only one shape is live at a time)
    # engine weights version last packed, per filter set: the packed filters sit at shape-independent offsets at the
    # head of the arena, so switching the batch shape does not invalidate them
    self.packed = {}

  def _shape(self, B, T):
    """Plan for this batch shape.  Ragged training produces a new (B, T) almost every step: plans (offsets + TMA
    descriptors, host-side only) are cached, the device arena is shared and only grows."""
    key = (B, T)
    sh = self.shapes.get(key)
    if sh is None:
      if len(self.shapes) >= MAX_CACHED_SHAPES:
        old = next(iter(self.shapes))
        self.shapes.pop(old).close()
      sh = _Shape(self.engine, B, T)
      if self.arena is None or self.arena.numel() < sh.nbytes + 1024:
        for other in self.shapes.values():
          other.close()
        self.shapes.clear()
        self.packed = {}
        self.arena = None                                      # release before growing
        self.arena = torch.zeros((int(sh.nbytes * 1.1) + 1024,), dtype=torch.uint8, device=self.engine.device)
      sh.bind(self.engine, self.arena)
      self.shapes[key] = sh
      if getattr(self, 'timing', False):
        check(lib().st_plan_set_timing(sh.handle, 1))
    return sh

  def set_timing(self, enable):
    """CUDA-event pairs around every tensor-core launch (recorded on the launch stream, inside the steps)."""
    self.timing = bool(enable)
    for sh in self.shapes.values():
      check(lib().st_plan_set_timing(sh.handle, int(self.timing)))

  def read_timings(self):
    """-> list of (kernel name, layer, algorithmic flops, ms); host-synchronous, clears the native log."""
    out = []
    names = {0: 'tc_conv_kernel', 1: 'tc_conv_kernel', 2: 'tc_wgrad_kernel'}
    tags = {0: 'fwd', 1: 'dgrad', 2: 'wgrad'}
    for sh in self.shapes.values():
      cap = 8192
      kind = (ctypes.c_int * cap)(); layer = (ctypes.c_int * cap)()
      flops = (ctypes.c_double * cap)(); ms = (ctypes.c_float * cap)()
      n = lib().st_plan_read_timings(sh.handle, kind, layer, flops, ms, cap)
      if n < 0:
        check(n)
      out += [(names[kind[i]], tags[kind[i]], layer[i], flops[i], ms[i]) for i in range(n)]
    return out

  def _launch_count(self, sh, before):
    self.engine.launches += lib().st_plan_launches(sh.handle) - before

  def _pack(self, sh):
    if self.packed.get(sh.filter_set) != self.engine._weights_version:
      before = lib().st_plan_launches(sh.handle)
      check(lib().st_plan_pack_weights(sh.handle, stream_ptr()))
      # a plan packs the layers common to every set plus its own layer-8 variant
      if self.packed.get('common') != self.engine._weights_version:
        self.packed = {'common': self.engine._weights_version}
      self.packed[sh.filter_set] = self.engine._weights_version
      self._launch_count(sh, before)

  def forward(self, inputs, keep_activations=False):
    eng = self.engine
    B, T, _ = inputs.shape
    sh = self._shape(B, T)
    self._pack(sh)
    before = lib().st_plan_launches(sh.handle)
    check(lib().st_plan_forward(sh.handle, ptr(inputs), stream_ptr()))
    self._launch_count(sh, before)
    self._last = sh
    return sh.logits_tm

  def activation(self, layer):
    """fp32 copy of layer `layer`'s output [B,T',Cout] from the last forward (tests)."""
    sh = self._last
    k, s, cin, cout, relu = self.engine.layers[layer]
    out = torch.empty((sh.B, sh.To, cout), dtype=torch.float32, device=self.engine.device)
    check(lib().st_plan_get_activation(sh.handle, layer, ptr(out), stream_ptr()))
    return out

  def train_step(self, inputs, sequence_lengths, labels, learning_rate, max_gradient_norm=5.0, decode=False):
    eng = self.engine
    B, T, _ = inputs.shape
    ctc_len = np.asarray(sequence_lengths, dtype=np.int32) // 2
    # the forward kernels are enqueued FIRST: flattening / validating the labels and their (pinned, asynchronous)
    # upload are host work that then runs underneath them instead of in front of the step
    # the flat gradient buffer is zeroed on the plan's side stream underneath the forward pass
    check(lib().st_plan_prepare_backward(self._shape(B, T).handle, stream_ptr()))
    logits = self.forward(inputs.contiguous(), keep_activations=True)
    batch = ops.CTCBatch(labels, ctc_len, -(-T // 2), eng.num_classes, eng.device)
    sh = self._last
    scale = 1.0 / (B * eng.world_size)
    loss, _ = ops.ctc_loss(batch, logits, want_grad=False, grad_scale=scale, grad_planes=sh.dlogits_planes)
    eng.launches += 3
    from .engine import BatchMean
    out = {'loss': loss, 'avg_loss': BatchMean(loss, early=True), 'decoded': None, 'logits': logits}
    if decode:
      out['decoded'], out['neg_sum_logits'] = ops.ctc_greedy_decoder(logits, ctc_len)
      eng.launches += 1
    before = lib().st_plan_launches(sh.handle)
    if True:
      if eng.world_size > 1:
        # layers 10..8 hold 82 % of the gradient bytes and finish first: their allreduce overlaps layers 7..0
        split = eng.layout.w_off[8]
        check(lib().st_plan_backward_range(sh.handle, 10, 8, stream_ptr()))
        handles = eng.allreduce_gradients(async_ranges=[(split, eng.grads.numel())])
        if DP_RESERVE_SMS > 0:       # optionally leave room for the allreduce kernel's CTAs (st_plan_reserve_sms)
          check(lib().st_plan_reserve_sms(sh.handle, DP_RESERVE_SMS))
        try:
          check(lib().st_plan_backward_range(sh.handle, 7, 0, stream_ptr()))
        finally:
          if DP_RESERVE_SMS > 0:
            check(lib().st_plan_reserve_sms(sh.handle, 0))
        handles += eng.allreduce_gradients(async_ranges=[(0, split)])
        for h in handles:
          h.wait()
      else:
        check(lib().st_plan_backward(sh.handle, stream_ptr()))
    self._launch_count(sh, before)
    eng.apply_gradients(learning_rate, max_gradient_norm, reduced=True)
    return out
