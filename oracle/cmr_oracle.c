/*
 * cmr_oracle.c - plain-C restatement of the integer-critical arithmetic of the
 * CMR-Agent geometric hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into or called by
 * the product (cmr_agent_b200); it exists so tests can check the CUDA kernels
 * bit-for-bit at full size in seconds, and so bench.py can report a CPU number.
 *
 * Each function cites the reference lines whose arithmetic it restates
 * (paths relative to /root/reference).  The floating-point evaluation ORDER is
 * the contract.  torch's CPU bmm with k=3 has TWO regimes
 * (measured, tests/test_oracle_vs_reference.py): when rows*cols*k < 400 (3x3 @ 3xn with
 * n <= 44: the 3x3 pose products of step/to_disentangled, and clouds of < 45 points) it is
 * the plain loop ((a0*x + a1*y) + a2*z), unfused; otherwise (MKL) it is the FMA chain
 * fma(a2,z, fma(a1,y, a0*x)).  The squared distances are unfused (dx*dx + dy*dy) + dz*dz.  Compile with -ffp-contract=off so the compiler
 * does not fuse or reorder anything that is not an explicit fmaf().
 *
 * Parity pinning: the reference ships no golden vectors; this file is pinned
 * against the real reference (torch CPU) by tests/test_oracle_vs_reference.py
 * in the build container and against tests/golden/ fixtures everywhere.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CMR_EXPORT __attribute__((visibility("default")))

static inline float dot3_fma(const float *a, float x, float y, float z) {
    /* one output element of a k=3 bmm with >= 45 columns (environment/environment.py:55,58,93,95) */
    return fmaf(a[2], z, fmaf(a[1], y, a[0] * x));
}
static inline float dot3_plain(const float *a, float x, float y, float z) {
    /* the same element when the product is small (rows*cols*k < 400): unfused, left to right */
    return (a[0] * x + a[1] * y) + a[2] * z;
}
static inline float dot3(const float *a, float x, float y, float z, int fused) {
    return fused ? dot3_fma(a, x, y, z) : dot3_plain(a, x, y, z);
}

/* environment/environment.py:54-72 (and :91-101) for every point of one cloud.
 * pc [3][N] channel-major, mean[3], RT[16] row-major 4x4, K[9] row-major.
 * fused = 1 when the bmm this column belongs to has >= 45 columns (see the header comment).
 * idx[j] = v*W+u (round-half-even) or H*W when out of frustum; in_cam[j] 0/1.  */
CMR_EXPORT void cmr_oracle_project(const float *pc, const float *mean, const float *RT, const float *K,
                                   int N, int H, int W, int fused, int32_t *idx, uint8_t *in_cam) {
    const float *px = pc, *py = pc + N, *pz = pc + 2 * (size_t)N;
    const float wmax = (float)(W - 1), hmax = (float)(H - 1);
    for (int j = 0; j < N; ++j) {
        float cx = px[j] - mean[0], cy = py[j] - mean[1], cz = pz[j] - mean[2];
        float X[3];
        for (int r = 0; r < 3; ++r) {
            float q = dot3(RT + 4 * r, cx, cy, cz, fused);
            X[r] = (q + mean[r]) + RT[4 * r + 3];
        }
        float U0 = dot3(K + 0, X[0], X[1], X[2], fused);
        float U1 = dot3(K + 3, X[0], X[1], X[2], fused);
        float U2 = dot3(K + 6, X[0], X[1], X[2], fused);
        float u = U0 / U2, v = U1 / U2;
        int ok = (u >= 0.0f) && (u <= wmax) && (v >= 0.0f) && (v <= hmax) && (U2 > 0.0f);
        in_cam[j] = (uint8_t)ok;
        if (ok) {
            int ui = (int)rintf(u), vi = (int)rintf(v); /* default rounding mode = half-to-even */
            idx[j] = vi * W + ui;
        } else {
            idx[j] = H * W;
        }
    }
}

/* environment/environment.py:74-82 with torch_scatter.scatter_mean semantics (oracle/shims.py):
 * sequential sum in point order over the predicted-overlap points, count clamped to >= 1.
 * feat [C][N] channel-major, overlap[N] 0/1, idx[N] from cmr_oracle_project; out [C][P]. */
CMR_EXPORT void cmr_oracle_scatter_mean(const float *feat, const uint8_t *overlap, const int32_t *idx,
                                        int N, int C, int P, float *out) {
    float *cnt = (float *)calloc((size_t)P, sizeof(float));
    memset(out, 0, sizeof(float) * (size_t)C * P);
    for (int j = 0; j < N; ++j)
        if (overlap[j] && idx[j] < P) cnt[idx[j]] += 1.0f;
    for (int c = 0; c < C; ++c) {
        const float *f = feat + (size_t)c * N;
        float *o = out + (size_t)c * P;
        for (int j = 0; j < N; ++j)
            if (overlap[j] && idx[j] < P) o[idx[j]] += f[j];
        for (int p = 0; p < P; ++p) o[p] = o[p] / (cnt[p] < 1.0f ? 1.0f : cnt[p]);
    }
    free(cnt);
}

/* environment/environment.py:284-290: per-episode mean over masked points of |a-b|^2, where
 * b = pc - mean (shipped) or the disentangled transform (intended, the commented line :273).
 * Accumulated in double (the reference's fp32 cascade sum is within 1e-6 of this). */
CMR_EXPORT double cmr_oracle_p2p(const float *target, const float *pc, const uint8_t *mask, const float *mean,
                                 const float *RT, int N, int intended) {
    double acc = 0.0;
    long cnt = 0;
    for (int j = 0; j < N; ++j) {
        if (!mask[j]) continue;
        float c[3] = {pc[j] - mean[0], pc[N + j] - mean[1], pc[2 * (size_t)N + j] - mean[2]};
        float b[3];
        for (int r = 0; r < 3; ++r)
            b[r] = intended ? (dot3(RT + 4 * r, c[0], c[1], c[2], N >= 45) + mean[r]) + RT[4 * r + 3] : c[r];
        float d0 = target[j] - b[0], d1 = target[N + j] - b[1], d2 = target[2 * (size_t)N + j] - b[2];
        float s = (d0 * d0 + d1 * d1) + d2 * d2;
        acc += (double)s;
        ++cnt;
    }
    return cnt ? acc / (double)cnt : NAN;
}

static inline float sqdist(const float *a, const float *b) {
    /* models/pointnet_util.py:33,67: sum((a-b)**2, -1), unfused, left to right */
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return (dx * dx + dy * dy) + dz * dz;
}

/* models/pointnet_util.py:50-70 for one cloud. xyz [N][3]; out[npoint]. */
CMR_EXPORT void cmr_oracle_fps(const float *xyz, int N, int npoint, int64_t start, int64_t *out) {
    float *dist = (float *)malloc(sizeof(float) * (size_t)N);
    for (int j = 0; j < N; ++j) dist[j] = 1e10f;
    int64_t far = start;
    for (int i = 0; i < npoint; ++i) {
        out[i] = far;
        const float *c = xyz + 3 * far;
        float best = -1.0f;
        int64_t arg = 0;
        for (int j = 0; j < N; ++j) {
            float d = sqdist(xyz + 3 * (size_t)j, c);
            if (d < dist[j]) dist[j] = d;          /* torch.min(distance, dist) */
            if (dist[j] > best) { best = dist[j]; arg = j; } /* torch.max -> first (lowest) index */
        }
        far = arg;
    }
    free(dist);
}

/* models/pointnet_util.py:115-116 with a STABLE order (distance, index); one cloud.
 * q [S][3], ref [N][3]; out [S][k]. Requires k <= N. */
CMR_EXPORT void cmr_oracle_knn(const float *q, const float *ref, int S, int N, int k, int64_t *out) {
    float *bd = (float *)malloc(sizeof(float) * (size_t)k);
    int64_t *bi = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
    for (int s = 0; s < S; ++s) {
        int n = 0;
        for (int j = 0; j < N; ++j) {
            float d = sqdist(q + 3 * (size_t)s, ref + 3 * (size_t)j);
            if (n == k && !(d < bd[k - 1])) continue; /* equal distance, larger index loses */
            int p = n < k ? n : k - 1;
            while (p > 0 && d < bd[p - 1]) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
            bd[p] = d; bi[p] = j;
            if (n < k) ++n;
        }
        memcpy(out + (size_t)s * k, bi, sizeof(int64_t) * (size_t)k);
    }
    free(bd); free(bi);
}

/* models/pointnet_util.py:73-93 for one cloud; r2 is float32(radius**2 evaluated in double). */
CMR_EXPORT void cmr_oracle_ball(const float *q, const float *ref, int S, int N, float r2, int nsample,
                                int64_t *out) {
    for (int s = 0; s < S; ++s) {
        int64_t *o = out + (size_t)s * nsample;
        int n = 0;
        for (int j = 0; j < N && n < nsample; ++j) {
            float d = sqdist(q + 3 * (size_t)s, ref + 3 * (size_t)j);
            if (!(d > r2)) o[n++] = j;
        }
        int64_t fill = n ? o[0] : N;
        for (; n < nsample; ++n) o[n] = fill;
    }
}

/* models/pointnet_util.py:19-33 dense matrix for one cloud: out [S][N]. */
CMR_EXPORT void cmr_oracle_sqdist(const float *q, const float *ref, int S, int N, float *out) {
    for (int s = 0; s < S; ++s)
        for (int j = 0; j < N; ++j) out[(size_t)s * N + j] = sqdist(q + 3 * (size_t)s, ref + 3 * (size_t)j);
}

/* environment/environment.py:204 for one pose: R <- Rnew @ R (3x3 @ 3x3 = the plain loop), t += move.
 * pose row-major 4x4 in place. */
CMR_EXPORT void cmr_oracle_apply_step(float *pose, const float *Rnew, const float *move_t) {
    float R[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R[3 * r + c] = dot3_plain(Rnew + 3 * r, pose[c], pose[4 + c], pose[8 + c]);
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) pose[4 * r + c] = R[3 * r + c];
        pose[4 * r + 3] = pose[4 * r + 3] + move_t[r];
    }
}

/* environment/environment.py:231-232: (Rx @ Ry) @ Rz, each 3x3 product the plain loop. */
CMR_EXPORT void cmr_oracle_compose_xyz(const float *Rx, const float *Ry, const float *Rz, float *out) {
    float A[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) A[3 * r + c] = dot3_plain(Rx + 3 * r, Ry[c], Ry[3 + c], Ry[6 + c]);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out[3 * r + c] = dot3_plain(A + 3 * r, Rz[c], Rz[3 + c], Rz[6 + c]);
}

/* environment/environment.py:19-20 for one pose: t <- (t - m) + R m  (3x3 @ 3x1 = the plain loop). */
CMR_EXPORT void cmr_oracle_to_disentangled(float *pose, const float *mean) {
    float t[3];
    for (int r = 0; r < 3; ++r) t[r] = (pose[4 * r + 3] - mean[r]) + dot3_plain(pose + 4 * r, mean[0], mean[1], mean[2]);
    for (int r = 0; r < 3; ++r) pose[4 * r + 3] = t[r];
}
