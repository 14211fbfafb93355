"""Drop-in for the reference's ``pre/time_test.py``: the for-loop entry point and the numba entry
point timed on one frame (pre/time_test.py:13-30).

The reference script cannot run as shipped: its paths are empty strings (:9-10), ``read_data``
is never called, and it hands ``cal_tsdf_cuda`` a dict with the key ``'depth'`` (:18-21) where
that function reads ``s['data']`` (pre/tsdf_numba.py:132-133).  Here ``read_data(raw, pc)``
takes the two paths (defaulting to ``RAW_dir`` / ``PC_dir``), passes each entry point the key
it reads, prints the reference's two lines and returns the two results.
"""
import time

import numpy as np

from tsdf_for import tsdf_f
from tsdf_numba import cal_tsdf_cuda

RAW_dir = ''
PC_dir = ''


def read_data(raw=None, pc_file=None):
    with open(raw or RAW_dir, 'rb') as f:
        header = np.fromfile(f, dtype=np.int32, count=6)
        depth = np.fromfile(f, dtype=np.float32)
    data = {'header': header, 'depth': depth, 'data': depth}
    pcs = np.load(pc_file or PC_dir)
    pc = pcs[0]
    t1 = time.time()
    tsdf1, max_l1, mid_p1 = tsdf_f(data, pc)
    t2 = time.time()
    tsdf2, max_l2, mid_p2 = cal_tsdf_cuda(data)
    t3 = time.time()
    print('for loop time:%.4f' % (t2 - t1))
    print('numba time:%.4f' % (t3 - t2))
    return (tsdf1, max_l1, mid_p1), (tsdf2, max_l2, mid_p2)


if __name__ == "__main__":
    import sys
    read_data(*sys.argv[1:3])
