"""Scheduling arithmetic of the persistent tensor-core kernels, restated in Python and checked exhaustively on the CPU.

tc_wgrad_kernel (csrc/conv_tc.cu) assigns (tile, K range) work items to CTAs -- or, on CTA pairs, tile PAIRS to
clusters -- with closed-form index arithmetic: whole waves, a K-sliced tail, or forced K slices.  A hole or an overlap
in that arithmetic would silently lose or double-count part of a filter gradient for SOME batch shape (ragged training
produces a new shape almost every step; the GPU suite sees a handful).  These tests mirror the device code line by
line (`item`, `my_items`) and the host helpers that must agree with it (`wgrad_accumulates`: does a launch accumulate
into its output, i.e. must the output be zeroed first; `wgrad_best_split`), over every shape class of the network.
"""
import itertools

import pytest


def wgrad_items(num_tiles_total, total_iters, grid, pair, force_split):
  """-> list over CTAs of [(tile of this CTA, q0, q1), ...] exactly as tc_wgrad_kernel computes them."""
  units = num_tiles_total >> (1 if pair else 0)
  G = grid >> (1 if pair else 0)
  full_waves = units // G
  tail = units - full_waves * G
  cap = total_iters // 4 if total_iters // 4 > 0 else 1
  tail_split = max(1, min(G // tail, cap)) if tail > 0 else 1
  fsplit = force_split if force_split > 1 else 0
  forced_items = fsplit * units
  out = []
  for block in range(grid):
    rank = block & 1 if pair else 0
    u = block >> 1 if pair else block
    if fsplit:
      n = (forced_items - u + G - 1) // G if u < forced_items else 0
    else:
      n = full_waves + (1 if u < tail * tail_split else 0)
    items = []
    for i in range(n):
      if fsplit:
        g = i * G + u
        sl = g // units
        tile = g - sl * units
        q0, q1 = total_iters * sl // fsplit, total_iters * (sl + 1) // fsplit
      elif i < full_waves:
        tile, q0, q1 = i * G + u, 0, total_iters
      else:
        sl = u // tail
        tile = full_waves * G + u % tail
        q0, q1 = total_iters * sl // tail_split, total_iters * (sl + 1) // tail_split
      if pair:
        tile = 2 * tile + rank
      items.append((tile, q0, q1))
    out.append(items)
  return out, tail_split if not fsplit else fsplit


def host_accumulates(num_tiles, total_iters, sms, pair):          # tc::wgrad_accumulates
  G = sms // 2 if pair else sms
  if pair:
    num_tiles //= 2
  tail = num_tiles % G
  if tail == 0:
    return False
  cap = total_iters // 4 if total_iters // 4 > 0 else 1
  return min(G // tail, cap) > 1


def host_best_split(num_tiles, total_iters, sms, pair):           # tc::wgrad_best_split
  G = sms // 2 if pair else sms
  if pair:
    num_tiles //= 2
  if num_tiles >= G:
    return 1
  best, best_cost = 1, ((num_tiles + G - 1) // G) * total_iters
  s = 2
  while s <= 8 and s * 4 <= total_iters:
    cost = ((num_tiles * s + G - 1) // G) * ((total_iters + s - 1) // s + 4)
    if cost < best_cost:
      best_cost, best = cost, s
    s += 1
  return best


def check_cover(num_tiles_total, total_iters, grid, pair, force_split):
  per_cta, split = wgrad_items(num_tiles_total, total_iters, grid, pair, force_split)
  covered = {}
  for block, items in enumerate(per_cta):
    for tile, q0, q1 in items:
      assert 0 <= tile < num_tiles_total and 0 <= q0 < q1 <= total_iters, (tile, q0, q1)
      covered.setdefault(tile, []).append((q0, q1))
  assert sorted(covered) == list(range(num_tiles_total)), 'a tile is never processed'
  sliced = False
  for tile, ranges in covered.items():
    ranges.sort()
    assert ranges[0][0] == 0 and ranges[-1][1] == total_iters, (tile, ranges)
    assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:])), (tile, ranges)     # no hole, no overlap
    sliced = sliced or len(ranges) > 1
  if pair:      # both CTAs of a cluster always work on the SAME K range of tiles 2u and 2u+1 (they share one MMA)
    for even, odd in zip(per_cta[0::2], per_cta[1::2]):
      assert [(t + 1, a, b) for t, a, b in even] == odd
  return sliced


# (taps, Cin tiles of 128, Cout tiles of 256) per filter-gradient launch of the network, times the problems in it
LAUNCHES = {'L0': (48, 1, 1, 1), 'L1-7 merged': (7, 2, 1, 7), 'L1 alone': (7, 2, 1, 1), 'L8 direct': (32, 2, 8, 1),
            'L8 one level': (16, 2, 8, 3), 'L8 two levels': (8, 2, 8, 9), 'L9': (1, 16, 8, 1)}


@pytest.mark.parametrize('name', sorted(LAUNCHES))
@pytest.mark.parametrize('sms', [148, 132, 140])                   # all SMs / SMs left by st_plan_reserve_sms
def test_wgrad_work_items_cover_every_tile_and_k_range_exactly_once(name, sms):
  taps, m_tiles, n_tiles, problems = LAUNCHES[name]
  tiles = problems * taps * m_tiles * n_tiles
  for batch, t_chunks in itertools.product([1, 2, 3, 4, 16, 32, 64, 256], [1, 2, 3, 8, 24]):
    total_iters = batch * t_chunks
    for pair in ([False, True] if (taps * m_tiles) % 2 == 0 else [False]):
      grid = 2 * (sms // 2) if pair else sms
      sliced = check_cover(tiles, total_iters, grid, pair, 0)
      # the host zeroes an output exactly when the kernel accumulates into it
      assert sliced == host_accumulates(tiles, total_iters, sms, pair), (name, batch, t_chunks, pair)
      forced = host_best_split(tiles, total_iters, sms, pair)
      if forced > 1:
        assert check_cover(tiles, total_iters, grid, pair, forced)


def test_pairing_needs_an_even_number_of_tap_and_cin_tile_combinations():
  """tc::wgrad_pair: tiles 2u and 2u+1 share their dZ tile only if they never straddle an n-tile boundary, i.e. if
  taps * m_tiles (the tiles per n tile) is even -- true for every layer of the network."""
  for name, (taps, m_tiles, n_tiles, problems) in LAUNCHES.items():
    per_n = taps * m_tiles
    assert per_n % 2 == 0, name
    for tile in range(0, problems * per_n * n_tiles, 2):
      assert tile // per_n == (tile + 1) // per_n                  # same (problem, n tile)
