"""Builds speecht_b200/libspeecht_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m speecht_b200.build [--force] [--verbose]

The shared library is the product's only native artefact; it is git-ignored but travels with the gpurun snapshot.
cudart is linked statically and the CUDA driver API (cuTensorMapEncodeTiled) is resolved at run time through
cudaGetDriverEntryPoint, so the .so loads -- and its symbols can be checked -- on a box without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ_DIR = os.path.join(HERE, 'build')
LIB_PATH = os.path.join(HERE, 'libspeecht_b200.so')
HEADER = os.path.join(os.path.dirname(HERE), 'include', 'speecht_b200.h')

SOURCES = ['st_api.cu', 'decode.cu', 'ctc.cu', 'conv_f32.cu', 'optim.cu', 'melspec.cu', 'conv_tc.cu', 'w2l_plan.cu',
           'flac_host.cu']

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']


def _nvcc():
  for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
      return cand
  return 'nvcc'


def _stale(target, deps):
  if not os.path.exists(target):
    return True
  t = os.path.getmtime(target)
  return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, report=None):
  """Compile what is stale (everything with force=True) and link.  `report`, when a list, receives the names of
  the sources that were actually recompiled and 'link' when the shared library was relinked."""
  os.makedirs(OBJ_DIR, exist_ok=True)
  nvcc = _nvcc()
  headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))] + [HEADER]
  sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
  jobs = []
  for src in sources:
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ_DIR, src.replace('.cu', '.o'))
    if force or _stale(obj, [path] + headers):
      jobs.append((path, obj))

  def compile_one(job):
    path, obj = job
    cmd = [nvcc] + NVCC_FLAGS + ['-c', path, '-o', obj]
    if verbose:
      print(' '.join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
      raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (path, r.stdout, r.stderr))
    return r.stderr

  with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
    for msg in pool.map(compile_one, jobs):
      if verbose and msg:
        print(msg)
  if report is not None:
    report.extend(os.path.basename(path) for path, _obj in jobs)
  objs = [os.path.join(OBJ_DIR, s.replace('.cu', '.o')) for s in sources]
  if force or jobs or _stale(LIB_PATH, objs):
    cmd = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-cudart', 'static', '-gencode', 'arch=compute_100a,code=sm_100a']
    if verbose:
      print(' '.join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
      raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    if report is not None:
      report.append('link')
  return LIB_PATH


if __name__ == '__main__':
  p = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
  print(p)
