/*
 * CPU ORACLE in C (test infrastructure, not product code).
 *
 * Plain-C restatement of the reference's depth-crop -> projective directional TSDF
 * path, statement for statement, in IEEE binary32 with no FMA contraction (compile
 * with -ffp-contract=off; x86-64 SSE evaluates float expressions in float).  It is
 * used (a) by tests/ to check the CUDA path on batches too large for the NumPy oracle,
 * and (b) by bench.py as the CPU baseline ("port"), threaded over frames with OpenMP.
 * It is itself pinned to the NumPy oracle and to the golden vectors made from the
 * unmodified reference (tests/test_oracle_c.py).  The product library never links it.
 *
 * Line citations are into /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SENTINEL 99999.0f /* pre/tsdf_numba.py:88-89 */

/* pre/tsdf_numba.py:83-96 (per pixel, Python-2 row index) + :98-116,140-141 (reduction);
 * the sentinels always take part (see oracle/dtsdf_oracle.py: bbox_from_pixels). */
void oracle_bbox_from_pixels(const int32_t* header, const float* depth, float focal, int cx, int cy,
                             float* min_p, float* max_p) {
    const int l = header[2], t = header[3], r = header[4], b = header[5];
    const int bw = r - l, bh = b - t;
    float mn[3] = {SENTINEL, SENTINEL, SENTINEL}, mx[3] = {-SENTINEL, -SENTINEL, -SENTINEL};
    const long n = (bw > 0 && bh > 0) ? (long)bw * bh : 0;
    for (long pos = 0; pos < n; ++pos) {
        const int x = (int)(pos % bw) + l;
        const int y = (int)(pos / bw) + t;
        float cam_z = depth[pos];
        if (fabsf(cam_z) < 1.0f) continue;
        const float q = cam_z / focal;
        const float cam_x = q * (float)(x - cx);
        const float cam_y = -q * (float)(y - cy);
        cam_z = -cam_z;
        if (cam_x < mn[0]) mn[0] = cam_x;
        if (cam_y < mn[1]) mn[1] = cam_y;
        if (cam_z < mn[2]) mn[2] = cam_z;
        if (cam_x > mx[0]) mx[0] = cam_x;
        if (cam_y > mx[1]) mx[1] = cam_y;
        if (cam_z > mx[2]) mx[2] = cam_z;
    }
    memcpy(min_p, mn, sizeof(mn));
    memcpy(max_p, mx, sizeof(mx));
}

/* pre/process.py:91-96 == pre/tsdf_for.py:11-16 == pre/tsdf_numba.py:142-147 */
void oracle_grid_params(const float* pmax, const float* pmin, int R, float trunc_mul, float* vox_ori,
                        float* voxel_len, float* trunc, float* max_l, float* mid) {
    float len_e[3];
    for (int c = 0; c < 3; ++c) {
        mid[c] = (pmax[c] + pmin[c]) / 2.0f;
        len_e[c] = pmax[c] - pmin[c];
    }
    float ml = len_e[0];
    if (len_e[1] > ml) ml = len_e[1];
    if (len_e[2] > ml) ml = len_e[2];
    const float vl = ml / (float)R;
    *max_l = ml;
    *voxel_len = vl;
    *trunc = vl * trunc_mul;
    for (int c = 0; c < 3; ++c) vox_ori[c] = mid[c] - ml / 2.0f + vl / 2.0f;
}

/* pre/process.py:140-198 == pre/tsdf_for.py:62-120 (layout 0, out[c][x][y][z]) and
 * pre/tsdf_numba.py:15-72 (layout 1, out[c][z][y][x]). */
void oracle_tsdf_cal(const int32_t* header, const float* depth, const float* vox_ori, float voxel_len,
                     float truncation, int R, int layout, float focal, int cx, int cy, float* out) {
    const int bb_left = header[2], bb_top = header[3], bb_right = header[4], bb_bottom = header[5];
    const int bb_width = bb_right - bb_left;
    const long ch = (long)R * R * R;
    memset(out, 0, sizeof(float) * 3 * ch);
    for (int x = 0; x < R; ++x) {
        for (int y = 0; y < R; ++y) {
            for (int z = 0; z < R; ++z) {
                /* voxel center (:150-152) */
                const float voxel_x = vox_ori[0] + (float)x * voxel_len;
                const float voxel_y = vox_ori[1] + (float)y * voxel_len;
                const float voxel_z = vox_ori[2] + (float)z * voxel_len;
                /* voxel center in image frame (:155-157); int() truncates toward zero */
                const float coeff = -focal / voxel_z;
                const float fpx = voxel_x * coeff + (float)cx;
                const float fpy = -voxel_y * coeff + (float)cy;
                if (!(fabsf(fpx) < 2e9f) || !(fabsf(fpy) < 2e9f)) continue; /* non-finite: outside */
                const int pixel_x = (int)fpx;
                const int pixel_y = (int)fpy;
                if (pixel_x < bb_left || pixel_x >= bb_right || pixel_y < bb_top || pixel_y >= bb_bottom)
                    continue; /* :159 */
                const long idx = (long)(pixel_y - bb_top) * bb_width + pixel_x - bb_left;
                const float pixel_depth = depth[idx];
                if (fabsf(pixel_depth) < 1.0f) continue; /* :167 */
                /* closest surface point in world frame (:172-175) */
                const float coeff1 = pixel_depth / focal;
                const float world_x = (float)(pixel_x - cx) * coeff1;
                const float world_y = (float)(-(pixel_y - cy)) * coeff1;
                const float world_z = -pixel_depth;
                float tsdf_x = fabsf(voxel_x - world_x) / truncation;
                float tsdf_y = fabsf(voxel_y - world_y) / truncation;
                float tsdf_z = fabsf(voxel_z - world_z) / truncation;
                const float dis = sqrtf(tsdf_x * tsdf_x + tsdf_y * tsdf_y + tsdf_z * tsdf_z);
                if (dis > 1.0f) { tsdf_x = 1.0f; tsdf_y = 1.0f; tsdf_z = 1.0f; }
                if (1.0f < tsdf_x) tsdf_x = 1.0f; /* python min(t, 1), :187-189 */
                if (1.0f < tsdf_y) tsdf_y = 1.0f;
                if (1.0f < tsdf_z) tsdf_z = 1.0f;
                if (world_z > voxel_z) { tsdf_x = -tsdf_x; tsdf_y = -tsdf_y; tsdf_z = -tsdf_z; }
                const long o = layout == 0 ? ((long)x * R + y) * R + z : ((long)z * R + y) * R + x;
                out[o] = tsdf_x;
                out[ch + o] = tsdf_y;
                out[2 * ch + o] = tsdf_z;
            }
        }
    }
}

/* Whole path over a packed batch (same input format as include/dtsdf_b200.h).
 * bbox_mode 0 = pixels (tsdf_numba.cal_tsdf_cuda, pre/tsdf_numba.py:119-161),
 *           1 = aux[i] = {max xyz, min xyz} (tsdf_for.tsdf_f after max_min_point),
 *           2 = aux[i] = {vox_ori xyz, voxel_len, trunc} (tsdf_cal arguments).
 * nthreads <= 0 uses the OpenMP default.  Returns the number of threads used. */
int oracle_dtsdf_batch(int n_frames, const int32_t* headers, const float* packed, const int64_t* offsets,
                       int R, int layout, int bbox_mode, const float* aux, float trunc_mul, float focal,
                       int cx, int cy, float* vol, float* max_l, float* mid_p, int nthreads) {
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
    {
#pragma omp single
        used = omp_get_num_threads();
    }
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int f = 0; f < n_frames; ++f) {
        const int32_t* h = headers + 6 * (long)f;
        const float* depth = packed + offsets[f];
        float vox_ori[3], voxel_len, trunc, ml = 0.0f, mid[3] = {0.0f, 0.0f, 0.0f};
        if (bbox_mode == 2) {
            const float* g = aux + 5 * (long)f;
            vox_ori[0] = g[0]; vox_ori[1] = g[1]; vox_ori[2] = g[2];
            voxel_len = g[3]; trunc = g[4];
        } else {
            float pmin[3], pmax[3];
            if (bbox_mode == 0) {
                oracle_bbox_from_pixels(h, depth, focal, cx, cy, pmin, pmax);
            } else {
                const float* g = aux + 6 * (long)f;
                pmax[0] = g[0]; pmax[1] = g[1]; pmax[2] = g[2];
                pmin[0] = g[3]; pmin[1] = g[4]; pmin[2] = g[5];
            }
            oracle_grid_params(pmax, pmin, R, trunc_mul, vox_ori, &voxel_len, &trunc, &ml, mid);
            if (max_l) max_l[f] = ml;
            if (mid_p) { mid_p[3 * (long)f] = mid[0]; mid_p[3 * (long)f + 1] = mid[1]; mid_p[3 * (long)f + 2] = mid[2]; }
        }
        const int bad = (h[4] - h[2] <= 0) || (h[5] - h[3] <= 0);
        float* out = vol + (long)f * 3 * R * R * R;
        if (bad) memset(out, 0, sizeof(float) * 3 * (long)R * R * R);
        else oracle_tsdf_cal(h, depth, vox_ori, voxel_len, trunc, R, layout, focal, cx, cy, out);
    }
    return used;
}

/* ------------------------------------------------------------------------------------------------
 * Augmented D-TSDF.  PARITY UNPINNED: the reference's data_aug raises
 * AxisError (pre/process.py:252) and the cut_version variant voxelises the un-augmented depth map
 * (cut_version/pre/process.py:207), so there is no reference behaviour to pin to.  This function IS
 * the specification the CUDA kernel (dtsdf_voxelize_aug_batch) is checked against, bit for bit:
 *
 *   p' = T(p; A, c) = (p - c) . A + c     row vector times matrix, as show/visualize.py:106-114
 *                                          (A = S . R: stretch then rotation), about the centre c
 *   each component: t = dx*A[0][k]; t = fmaf(dy, A[1][k], t); t = fmaf(dz, A[2][k], t); q_k = t + c_k
 *
 *   1. bounding box over the AUGMENTED points of all valid pixels (|d| >= 1; back-projection as
 *      pre/tsdf_numba.py:91-96), sentinels as pre/tsdf_numba.py:88-89;
 *   2. grid parameters from it, pre/process.py:91-96;
 *   3. per voxel centre v' = (v'_x[x], v'_y[y], v'_z[z]) of that grid (augmented space): back in the camera
 *      frame v = (v' - c) . Ainv + c, evaluated SEPARABLY - one table per grid axis, three components each,
 *      and a fixed order of the two additions (round 2: the map is linear in the three grid indices, so a
 *      per-axis table sum costs 3-4 additions per voxel where the row-vector product cost 15 operations):
 *          Tx[x][k] = (v'_x[x] - c_x) * Ainv[0][k]
 *          Ty[y][k] = (v'_y[y] - c_y) * Ainv[1][k]
 *          Tz[z][k] = fmaf(v'_z[z] - c_z, Ainv[2][k], c_k)
 *          v_k      = Tx[x][k] + (Ty[y][k] + Tz[z][k])
 *      then: project v (pre/process.py:155-157), look the depth up (:159-167), back-project the pixel to the
 *      surface point w (:172-175); the distance vector of the augmented frame is the camera-frame difference
 *      under the linear part of T - v' - w' = (v - w) . A, the translation cancels - with the products in T's
 *      own order:
 *          e_k = (v_x - w_x)*A[0][k];  e_k = fmaf(v_y - w_y, A[1][k], e_k);  e_k = fmaf(v_z - w_z, A[2][k], e_k)
 *      (round 2, second step: no second transformed point, no centre and no table of voxel centres in the
 *      per-voxel evaluation - 12 operations where T(w) and the subtraction from v' cost 21 and three table
 *      reads); per-axis distances |e_k| / trunc, Euclidean truncation and min(t,1) (:177-189); negative behind
 *      the surface as seen from the camera: w_z > v_z (:191-194, in the camera frame).
 * With A = Ainv = identity and any finite c the result equals oracle_dtsdf_batch with bbox_mode 0
 * whenever (p - c) + c == p exactly, in particular for c = 0 (the table sum then adds signed zeros to the one
 * non-zero term of each component).
 * aug[f] = { A[3][3], Ainv[3][3], c[3] } (21 floats, row-major matrices).
 */
static inline void xform(const float* p, const float* M, const float* c, float* q) {
    const float dx = p[0] - c[0], dy = p[1] - c[1], dz = p[2] - c[2];
    for (int k = 0; k < 3; ++k) {
        float t = dx * M[k];
        t = fmaf(dy, M[3 + k], t);
        t = fmaf(dz, M[6 + k], t);
        q[k] = t + c[k];
    }
}

void oracle_dtsdf_aug_frame(const int32_t* header, const float* depth, const float* aug, int R, int layout,
                            float trunc_mul, float focal, int cx, int cy, float* out, float* max_l_out,
                            float* mid_out) {
    const float *A = aug, *Ainv = aug + 9, *c = aug + 18;
    const int l = header[2], t = header[3], r = header[4], b = header[5];
    const int bw = r - l, bh = b - t;
    const long ch = (long)R * R * R;
    memset(out, 0, sizeof(float) * 3 * ch);
    float mn[3] = {SENTINEL, SENTINEL, SENTINEL}, mx[3] = {-SENTINEL, -SENTINEL, -SENTINEL};
    const long n = (bw > 0 && bh > 0) ? (long)bw * bh : 0;
    for (long pos = 0; pos < n; ++pos) {
        const float d = depth[pos];
        if (fabsf(d) < 1.0f) continue;
        const int x = (int)(pos % bw) + l, y = (int)(pos / bw) + t;
        const float q = d / focal;
        const float p[3] = {q * (float)(x - cx), -q * (float)(y - cy), -d};
        float pa[3];
        xform(p, A, c, pa);
        for (int k = 0; k < 3; ++k) {
            if (pa[k] < mn[k]) mn[k] = pa[k];
            if (pa[k] > mx[k]) mx[k] = pa[k];
        }
    }
    float vox_ori[3], voxel_len, trunc, ml, mid[3];
    oracle_grid_params(mx, mn, R, trunc_mul, vox_ori, &voxel_len, &trunc, &ml, mid);
    if (max_l_out) *max_l_out = ml;
    if (mid_out) { mid_out[0] = mid[0]; mid_out[1] = mid[1]; mid_out[2] = mid[2]; }
    if (n == 0) return;
    float Tx[64][3], Ty[64][3], Tz[64][3];
    for (int i = 0; i < R; ++i) {
        const float dx = (vox_ori[0] + (float)i * voxel_len) - c[0];
        const float dy = (vox_ori[1] + (float)i * voxel_len) - c[1];
        const float dz = (vox_ori[2] + (float)i * voxel_len) - c[2];
        for (int k = 0; k < 3; ++k) {
            Tx[i][k] = dx * Ainv[k];
            Ty[i][k] = dy * Ainv[3 + k];
            Tz[i][k] = fmaf(dz, Ainv[6 + k], c[k]);
        }
    }
    for (int x = 0; x < R; ++x)
        for (int y = 0; y < R; ++y)
            for (int z = 0; z < R; ++z) {
                float v[3];
                for (int k = 0; k < 3; ++k) v[k] = Tx[x][k] + (Ty[y][k] + Tz[z][k]);
                const float coeff = -focal / v[2];
                const float fpx = v[0] * coeff + (float)cx;
                const float fpy = -v[1] * coeff + (float)cy;
                if (!(fabsf(fpx) < 2e9f) || !(fabsf(fpy) < 2e9f)) continue;
                const int px = (int)fpx, py = (int)fpy;
                if (px < l || px >= r || py < t || py >= b) continue;
                const float d = depth[(long)(py - t) * bw + px - l];
                if (fabsf(d) < 1.0f) continue;
                const float c1 = d / focal;
                const float w[3] = {(float)(px - cx) * c1, (float)(-(py - cy)) * c1, -d};
                const float dlt[3] = {v[0] - w[0], v[1] - w[1], v[2] - w[2]};
                float e[3];
                for (int k = 0; k < 3; ++k) {
                    e[k] = dlt[0] * A[k];
                    e[k] = fmaf(dlt[1], A[3 + k], e[k]);
                    e[k] = fmaf(dlt[2], A[6 + k], e[k]);
                }
                float tx = fabsf(e[0]) / trunc;
                float ty = fabsf(e[1]) / trunc;
                float tz = fabsf(e[2]) / trunc;
                const float dis = sqrtf(tx * tx + ty * ty + tz * tz);
                if (dis > 1.0f) { tx = 1.0f; ty = 1.0f; tz = 1.0f; }
                if (1.0f < tx) tx = 1.0f;
                if (1.0f < ty) ty = 1.0f;
                if (1.0f < tz) tz = 1.0f;
                if (w[2] > v[2]) { tx = -tx; ty = -ty; tz = -tz; }
                const long o = layout == 0 ? ((long)x * R + y) * R + z : ((long)z * R + y) * R + x;
                out[o] = tx;
                out[ch + o] = ty;
                out[2 * ch + o] = tz;
            }
}

int oracle_dtsdf_aug_batch(int n_frames, const int32_t* headers, const float* packed, const int64_t* offsets,
                           const float* aug, int R, int layout, float trunc_mul, float focal, int cx, int cy,
                           float* vol, float* max_l, float* mid_p) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int f = 0; f < n_frames; ++f)
        oracle_dtsdf_aug_frame(headers + 6 * (long)f, packed + offsets[f], aug + 21 * (long)f, R, layout, trunc_mul,
                               focal, cx, cy, vol + (long)f * 3 * R * R * R, max_l ? max_l + f : 0,
                               mid_p ? mid_p + 3 * (long)f : 0);
    return 0;
}
