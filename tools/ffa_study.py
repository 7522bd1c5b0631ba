"""CPU study for the fast-FIR split of the 32-tap layer: checks the index algebra of the
three-convolution form against the direct 'SAME' correlation and compares their rounding error when every operand is
a bf16 hi+lo pair, products hi*hi + hi*lo + lo*hi, fp32 accumulation (numpy emulation of the bf16x3 tensor path; the
tensor core's own accumulate truncation is not modelled -- it adds the same ~2e-5 to both forms).

  python tools/ffa_study.py [seed]

Backward of the same form (checked against torch autograd on a small case by check_backward below): with
dA00[u] = dy[2u] - dy[2u-1], dA11[u] = dy[2u] - dy[2u+1], dS[u] = dy[2u+1] the data gradient is three half-rate
transposed 16-tap convolutions d_odd = dA00 (*) w0, d_even = dA11 (*) w1, d_xs = dS (*) ws with dx[2r+1] = d_odd[r] +
d_xs[r], dx[2r] = d_even[r] + d_xs[r], and the filter gradient three half-rate correlations dw[2j] = <odd, dA00>_j +
<xs, dS>_j, dw[2j+1] = <even, dA11>_j + <xs, dS>_j -- 75 % of the MMAs in every pass.
"""
import sys

import numpy as np

K, PAD_L = 32, 15


def bf16_round(x):
  """round-to-nearest-even to bfloat16, returned as float32"""
  u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
  r = ((u + 0x7fff + ((u >> 16) & 1)) >> 16) << 16
  return r.astype(np.uint32).view(np.float32)


def split(x):
  hi = bf16_round(x)
  lo = bf16_round(np.asarray(x, np.float32) - hi)
  return hi, lo


def mm3(a, b):
  """a [M,Kc] @ b [Kc,N] with bf16x3 operands, fp32 accumulation: main (hi*hi) + side (hi*lo + lo*hi)"""
  ah, al = split(a)
  bh, bl = split(b)
  main = ah @ bh
  side = ah @ bl + al @ bh
  return (main + side).astype(np.float32)


def corr(xp, w, rows, mm):
  """sum_j xp[u + j] @ w[j] for u in range(rows): xp [U, Cin], w [J, Cin, Cout]"""
  J = w.shape[0]
  out = None
  for j in range(J):
    term = mm(xp[j:j + rows], w[j])
    out = term if out is None else out + term
  return out


def main():
  seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
  rng = np.random.default_rng(seed)
  T, cin, cout = 501, 250, 192
  x = np.maximum(rng.standard_normal((T, cin)), 0).astype(np.float32)          # a ReLU output
  lim = np.sqrt(6.0 / (K * cin + K * 2000))
  w = rng.uniform(-lim, lim, size=(K, cin, cout)).astype(np.float32)
  xp = np.zeros((T + K, cin), np.float32)                                      # x'[n] = x[n - 15], zero outside
  xp[PAD_L:PAD_L + T] = x

  exact = corr(xp.astype(np.float64), w.astype(np.float64), T, lambda a, b: a @ b)
  direct = corr(xp, w, T, mm3)

  # fast-FIR form: even / odd phases of x' and of the taps
  x0, x1 = xp[0::2], xp[1::2]                                                   # x'_0[u] = x'[2u], x'_1[u] = x'[2u+1]
  w0, w1 = w[0::2], w[1::2]
  n_even, n_odd = (T + 1) // 2, T // 2
  x0s = x0[1:]                                                                  # x'_0[u + 1]
  xs = (x1[:x0s.shape[0]] + x0s).astype(np.float32)                             # x'_1[u] + x'_0[u+1], re-split inside mm3
  ws = (w0 + w1).astype(np.float32)

  def run(mm, dt):
    a00 = corr(x0.astype(dt), w0.astype(dt), n_even + 1, mm)                    # needed at u and u + 1
    a11 = corr(x1.astype(dt), w1.astype(dt), n_even, mm)
    s = corr(xs.astype(dt), ws.astype(dt), n_odd, mm)
    y = np.empty((T, cout), dt)
    y[0::2] = a00[:n_even] + a11[:n_even]
    y[1::2] = s - a11[:n_odd] - a00[1:n_odd + 1]
    return y

  # The same three convolutions as the plan would launch them: stride-1 'SAME'-style correlations with 16 taps over
  # row VIEWS of the unpadded activation x (TMA zero fill outside [0, rows)):
  #   A00[u] = sum_j odd [u + j - 8] w0[j]   odd [r] = x[2r + 1]   (pad_left 8)
  #   A11[u] = sum_j even[u + j - 7] w1[j]   even[r] = x[2r]       (pad_left 7)
  #   S[u]   = sum_j xs  [u + j - 7] ws[j]   xs  [r] = x[2r] + x[2r + 1]   (pad_left 7)
  def view_corr(v, wj, pad_left, rows):
    vp = np.zeros((rows + wj.shape[0] + pad_left + 1, v.shape[1]), np.float64)
    n = min(v.shape[0], vp.shape[0] - pad_left)
    vp[pad_left:pad_left + n] = v[:n]
    return corr(vp, wj.astype(np.float64), rows, lambda a, b: a @ b)

  x64 = x.astype(np.float64)
  odd, even = x64[1::2], x64[0::2]
  xs_v = even.copy()
  xs_v[:odd.shape[0]] += odd
  a00_v = view_corr(odd, w0, 8, n_even + 1)
  a11_v = view_corr(even, w1, 7, n_even)
  s_v = view_corr(xs_v, ws, 7, n_odd)
  y_v = np.empty((T, cout), np.float64)
  y_v[0::2] = a00_v[:n_even] + a11_v[:n_even]
  y_v[1::2] = s_v - a11_v[:n_odd] - a00_v[1:n_odd + 1]
  print('row-view formulation (odd/pad 8, even/pad 7, pair sums/pad 7) vs direct: max abs diff / max |y| = %.2e'
        % (np.max(np.abs(y_v - exact)) / np.max(np.abs(exact))))

  ffa64 = run(lambda a, b: a @ b, np.float64)
  ffa = run(mm3, np.float32)
  scale = np.max(np.abs(exact))
  print('index algebra (float64 fast-FIR vs direct):   max abs diff / max |y| = %.2e' % (np.max(np.abs(ffa64 - exact)) / scale))
  for name, y in (('direct   bf16x3', direct), ('fast-FIR bf16x3', ffa)):
    e = np.abs(y - exact)
    print('%s: max err / max|y| = %.2e   rms err / rms y = %.2e   even rows max %.2e   odd rows max %.2e'
          % (name, e.max() / scale, np.sqrt(np.mean(e ** 2)) / np.sqrt(np.mean(exact ** 2)), e[0::2].max() / scale,
             e[1::2].max() / scale))
  it_direct, it_ffa = K * 4, 3 * (K // 2) * 4
  print('pipeline iterations per 128 output rows x 256 channels: direct %d, fast-FIR %d per 2 x 128 rows -> %.0f %%'
        % (it_direct, it_ffa, 100.0 * it_ffa / (2 * it_direct)))


def check_backward():
  """dx and dw of the three-correlation form (docstring formulas) against autograd of the direct 32-tap correlation."""
  import torch
  torch.manual_seed(0)
  K, PAD = 32, 15
  T, cin, cout = 101, 6, 5
  x = torch.randn(T, cin, dtype=torch.float64, requires_grad=True)
  w = torch.randn(K, cin, cout, dtype=torch.float64, requires_grad=True)
  xp = torch.zeros(T + K, cin, dtype=torch.float64)
  xp = torch.cat([torch.zeros(PAD, cin, dtype=torch.float64), x, torch.zeros(K - PAD, cin, dtype=torch.float64)])
  y = sum(xp[k:k + T] @ w[k] for k in range(K))
  dy = torch.randn(T, cout, dtype=torch.float64)
  (y * dy).sum().backward()
  dx_ref, dw_ref = x.grad.numpy(), w.grad.numpy()
  xn, wn, dyn = x.detach().numpy(), w.detach().numpy(), dy.numpy()
  n_even, n_odd = (T + 1) // 2, T // 2
  def get(a, i):  # zero outside
      return a[i] if 0 <= i < a.shape[0] else np.zeros(a.shape[1])
  dye = dyn[0::2]; dyo = dyn[1::2]
  U = n_even + 1
  dA00 = np.stack([get(dye, u) - get(dyo, u - 1) for u in range(U)])
  dA11 = np.stack([get(dye, u) - get(dyo, u) for u in range(n_even)])
  dS = np.stack([get(dyo, u) for u in range(n_odd)])
  w0, w1 = wn[0::2], wn[1::2]; ws = w0 + w1
  odd, even = xn[1::2], xn[0::2]
  xs = even.copy(); xs[:odd.shape[0]] += odd
  # forward views: A00[u] = sum_j odd[u+j-8] w0[j]; A11[u] = sum_j even[u+j-7] w1[j]; S[u] = sum_j xs[u+j-7] ws[j]
  def tconv(dA, wj, pad, rows):   # d_view[r] = sum_j dA[r - j + pad] wj[j]^T
      out = np.zeros((rows, wj.shape[1]))
      for r in range(rows):
          for j in range(wj.shape[0]):
              u = r - j + pad
              if 0 <= u < dA.shape[0]:
                  out[r] += dA[u] @ wj[j].T
      return out
  d_odd = tconv(dA00, w0, 8, odd.shape[0]); d_even = tconv(dA11, w1, 7, even.shape[0]); d_xs = tconv(dS, ws, 7, even.shape[0])
  dx = np.zeros_like(xn)
  dx[1::2] = d_odd + d_xs[:odd.shape[0]]
  dx[0::2] = d_even + d_xs
  e_dx = np.abs(dx - dx_ref).max()
  def wcorr(v, dA, pad, J):       # <v, dA>_j = sum_u v[u + j - pad]^T dA[u]
      out = np.zeros((J, v.shape[1], dA.shape[1]))
      for j in range(J):
          for u in range(dA.shape[0]):
              r = u + j - pad
              if 0 <= r < v.shape[0]:
                  out[j] += np.outer(v[r], dA[u])
      return out
  c0 = wcorr(odd, dA00, 8, 16); c1 = wcorr(even, dA11, 7, 16); cs = wcorr(xs, dS, 7, 16)
  dw = np.zeros_like(wn); dw[0::2] = c0 + cs; dw[1::2] = c1 + cs
  print('backward formulas vs autograd (float64, T=101): max |dx err| = %.1e, max |dw err| = %.1e' % (e_dx, np.abs(dw - dw_ref).max()))



if __name__ == '__main__':
  main()
  check_backward()
