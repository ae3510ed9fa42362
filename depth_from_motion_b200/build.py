"""In-tree build of the C-ABI CUDA library (nvcc, sm_100a only)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'dfm_b200.cu')
OUT = os.path.join(HERE, 'libdfm_b200.so')
DEPS = [os.path.join(HERE, 'csrc', f) for f in
        ('dfm_b200.cu', 'common.cuh', 'simt_kernels.cuh', 'conv_tc.cuh', 'conv_tc_neck.cuh',
         'neck_api.inc', 'frustum_api.inc', 'frustum_kernels.cuh', 'pipeline_api.inc', 'bev_api.inc', 'tail_kernels.cuh', 'voxel_sample_api.inc', 'stereo_tail_api.inc', 'logits_tc.cuh')] + [os.path.join(HERE, '..', 'include', 'dfm_b200.h')]


def nvcc_path():
    return shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'


def up_to_date():
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False):
    if up_to_date() and not force:
        return OUT
    cmd = [nvcc_path(), '-gencode', 'arch=compute_100a,code=sm_100a', '-O3',
           '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC', '-shared',
           '-o', OUT, SRC]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
