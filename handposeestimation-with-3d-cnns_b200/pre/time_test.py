"""Drop-in for the reference's ``pre/time_test.py``: one frame through the for-loop entry point
(``tsdf_for.tsdf_f``) and through the numba entry point (``tsdf_numba.cal_tsdf_cuda``), with the
wall time of each printed the way the reference prints it (pre/time_test.py:13-30).

The reference script cannot run as shipped: both paths are empty strings (:9-10), ``read_data`` is
never called, and the frame dict it builds has the key ``'depth'`` (:18-21) while ``cal_tsdf_cuda``
reads ``s['data']`` (pre/tsdf_numba.py:132-133).  Here ``read_data`` takes the two paths as
arguments (falling back to the module constants), hands each entry point a record with the key it
reads, and returns both results besides printing.
"""
import sys
from timeit import default_timer

import numpy as np

import read_MSRA
import tsdf_for
import tsdf_numba

RAW_dir = ''    # one '%06d_depth.bin' file
PC_dir = ''     # a Point_Cloud/<gesture>.npy file; its first cloud is used


def _clock(fn, *args):
    start = default_timer()
    result = fn(*args)
    return result, default_timer() - start


def read_data(raw=None, pc_file=None):
    hdr, crop = read_MSRA.read_bin(raw if raw else RAW_dir)
    cloud = np.load(pc_file if pc_file else PC_dir)[0]
    loop_result, loop_s = _clock(tsdf_for.tsdf_f, {'header': hdr, 'depth': crop}, cloud)
    cuda_result, cuda_s = _clock(tsdf_numba.cal_tsdf_cuda, {'header': hdr, 'data': crop})
    print('for loop time:%.4f' % loop_s)
    print('numba time:%.4f' % cuda_s)
    return loop_result, cuda_result


if __name__ == "__main__":
    read_data(*sys.argv[1:3])
