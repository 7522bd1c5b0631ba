"""Index algebra of the NESTED fast-FIR split (two levels: a 32-tap correlation as nine quarter-rate 8-tap ones),
checked in float64 against the direct correlation, forward and backward (autograd).  General node, any pad parity:

  y[n] = sum_k v[n + k - P] h[k],  h0 = h[0::2], h1 = h[1::2], hs = h0 + h1, ve[r] = v[2r], vo[r] = v[2r+1]
  P = 2p+1:  X = vo (*) h0 (pad p+1),  Y = ve (*) h1 (pad p),  Z = (ve + vo) (*) hs (pad p)
  P = 2p  :  X = ve (*) h0 (pad p),    Y = vo (*) h1 (pad p),  Z = c (*) hs (pad p-1) with c[r] = v[2r-1] + v[2r]
  y[2u] = X[u] + Y[u],   y[2u+1] = Z[u] - Y[u] - X[u+1]

  python tools/ffa2_study.py
"""
import numpy as np
import torch


def direct(v, h, P, N):
  K = h.shape[0]
  vp = torch.zeros((N + K + P + 2, v.shape[1]), dtype=v.dtype)
  n = min(v.shape[0], vp.shape[0] - P)
  vp = torch.cat([torch.zeros((P, v.shape[1]), dtype=v.dtype), v[:n], torch.zeros((vp.shape[0] - P - n, v.shape[1]), dtype=v.dtype)])
  return sum(vp[k:k + N] @ h[k] for k in range(K))


def zpad(a, n):
  return a if a.shape[0] >= n else torch.cat([a, torch.zeros((n - a.shape[0],) + tuple(a.shape[1:]), dtype=a.dtype)])


def ffa(v, h, P, N, depth, count=None):
  if depth == 0:
    if count is not None:
      count.append((h.shape[0], N))
    return direct(v, h, P, N)
  h0, h1 = h[0::2], h[1::2]
  hs = h0 + h1
  ve, vo = v[0::2], v[1::2]
  n_even, n_odd = (N + 1) // 2, N // 2
  if P % 2 == 1:
    p = (P - 1) // 2
    L = max(ve.shape[0], vo.shape[0])
    a, pa, b, pb, c, pc = vo, p + 1, ve, p, zpad(ve, L) + zpad(vo, L), p
  else:
    p = P // 2
    assert p >= 1
    # c[r] = ve[r] + vo[r-1] = v[2r] + v[2r-1] (the pair straddles the even index; vo[-1] = 0), correlated with pad p-1
    L = max(ve.shape[0], vo.shape[0] + 1)
    vo_shift = torch.cat([torch.zeros((1, v.shape[1]), dtype=v.dtype), vo])
    a, pa, b, pb, c, pc = ve, p, vo, p, zpad(ve, L) + zpad(vo_shift, L), p - 1
  X = ffa(a, h0, pa, n_even + 1, depth - 1, count)
  Y = ffa(b, h1, pb, n_even, depth - 1, count)
  Z = ffa(c, hs, pc, max(n_odd, 1), depth - 1, count)
  y = torch.zeros((N, h.shape[2]), dtype=v.dtype)
  y[0::2] = X[:n_even] + Y[:n_even]
  y[1::2] = Z[:n_odd] - Y[:n_odd] - X[1:n_odd + 1]
  return y


def main():
  torch.manual_seed(0)
  for T in (501, 500, 51, 7):
    cin, cout, K, P = 5, 4, 32, 15
    x = torch.randn(T, cin, dtype=torch.float64, requires_grad=True)
    w = torch.randn(K, cin, cout, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(T, cout, dtype=torch.float64)
    y0 = direct(x, w, P, T)
    (y0 * dy).sum().backward()
    gx, gw = x.grad.clone(), w.grad.clone()
    for depth in (1, 2):
      x.grad = None; w.grad = None
      count = []
      y = ffa(x, w, P, T, depth, count)
      (y * dy).sum().backward()
      work = sum(k * -(-n // 128) for k, n in count)
      print('T=%3d depth %d: fwd %.1e  dx %.1e  dw %.1e   leaf problems %d (taps x rows: %s), tap-tiles %d vs direct %d'
            % (T, depth, (y - y0).abs().max(), (x.grad - gx).abs().max(), (w.grad - gw).abs().max(), len(count),
               sorted(set(count)), work, K * -(-T // 128)))


if __name__ == '__main__':
  main()
