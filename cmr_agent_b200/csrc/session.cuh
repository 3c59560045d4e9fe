// session.cuh - the rollout session: a registration rollout (Test_Agent.py:150-170 with scripted or precomputed
// actions) driven from HOST buffers by one native call per rollout, pipelined over `depth` slots so that the upload
// of rollout k+1 (copy stream) runs under the kernels of rollout k (compute stream).  Everything a slot needs lives
// in HBM for the session's lifetime (inputs, workspace, observations, results); results come back through pinned
// staging.  This is host-side runtime code: it owns streams, events and memory, and launches the same entry points
// the drop-in modules call - there is no second implementation of the path.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/cmr_b200.h"

namespace cmr {

// pc_mask arrives as int64 (dataset/KittiDataset.py:408); the reward reads `mask != 0` as bytes (environment.py:268)
__global__ void k_mask_to_u8(const long long *__restrict__ m, unsigned char *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = m[i] != 0;
}
__global__ void k_pose_identity(float *pose, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 16) pose[i] = ((i & 15) % 5 == 0) ? 1.f : 0.f;
}

struct SessionSlot {
    // device
    float *pc = nullptr, *feat = nullptr, *img_feat = nullptr, *K = nullptr, *target_pose = nullptr, *pc_in_cam = nullptr;
    unsigned char *overlap = nullptr, *mask_u8 = nullptr;
    long long *mask_i64 = nullptr, *a_r = nullptr, *a_t = nullptr;
    float *mean = nullptr, *pose = nullptr, *obs2d = nullptr, *obs3d = nullptr, *rew = nullptr, *dist = nullptr;
    void *ws = nullptr, *scratch = nullptr;
    // pinned host staging for the results
    float *h_rew = nullptr, *h_dist = nullptr, *h_pose = nullptr, *h_target = nullptr;
    cudaEvent_t uploaded = nullptr, done = nullptr, up_begin = nullptr, up_end = nullptr;
    bool in_flight = false;
    const float *feat_used = nullptr, *img_used = nullptr;   // device pointers the kernels read (own copy or the caller's)
};

}  // namespace cmr

struct cmr_session {
    cmr_session_config cfg;
    int depth = 0;
    cmr::SessionSlot *slots = nullptr;
    cudaStream_t copy = nullptr, compute = nullptr;
    float *rot_tab = nullptr, *t_tab = nullptr;   // device
    int nbins = 0;
    long long submitted = 0;
    double h2d_bytes = 0.0, h2d_seconds = 0.0;
    long long timed_uploads = 0;
    size_t bytes_per_upload = 0;
};
