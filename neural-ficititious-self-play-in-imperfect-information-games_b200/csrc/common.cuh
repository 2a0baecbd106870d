// Shared host-side plumbing of libnfsp_b200.so: the handle, error reporting, launch helpers.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/nfsp_b200.h"

struct nfsp_env_s {
    int rules;
    int device;
    int sm_count;
    int64_t n;
    uint64_t seed, game0, step;
    uint64_t *d_state;   // n packed words
    uint32_t *d_fsm;     // table image of the rule set's state-machine kernel (nfsp_fsm.cuh / legacy_fsm.cuh), built at create
    float *d_wpack;      // acting nets repacked for the kernels (act_kernels.cu), or nullptr
    uint32_t *d_work;    // dynamic work counter of the fused rollout
    void *d_wtc_wide;    // tensor-core operand image of layer 1, the four nets side by side along N (act_tc_kernels.cu)
    bool has_weights;
    // the tensor-core variants' weight images are built from the handle's own copy of the nets (behind the packed images
    // in d_wpack) the first time one of them runs after nfsp_act_set_weights: the default path pays for one pack launch
    bool tc_dirty = false, tq_dirty = false;
    bool st_dirty = false;    // the table of the nets' outputs per decision state (rollout_states.cu), behind d_wcopy in d_wpack
    float *d_states = nullptr;
    const float *d_wcopy = nullptr;
    // warp-specialised tcgen05 rollout (rollout_tq.cu)
    int w2_slot = -1;         // this handle's slot of the constant-bank image of the second layers, -1 = none yet
    void *d_w2img = nullptr;  // device copy of that image
    uint32_t *d_err = nullptr;  // protocol-error word of the rollout kernel (0 = none), see nfsp_env_kernel_error
    uint32_t tq_patience[2] = {1500u, 700u};  // cycles a partly filled tile of an average / best-response net may wait
};

// builds and uploads the table image of nfsp_step_fsm_kernel (env_kernels.cu); the handle's device is current
int nfsp_fsm_upload(nfsp_env_t h);

// builds the tensor-core operand image of the acting nets (act_tc_kernels.cu)
int nfsp_pack_tc_image(nfsp_env_t h, const float *d_weights, cudaStream_t st);
// brings the lazily built images of the tensor-core variants up to date on `st` (act_kernels.cu)
int nfsp_ensure_tc_images(nfsp_env_t h, bool tq, cudaStream_t st);
// second-layer image of the warp-specialised rollout: constant-bank slot of the handle (rollout_tq.cu)
int nfsp_tq_set_weights(nfsp_env_t h, const float *d_weights, cudaStream_t st);
void nfsp_tq_release(nfsp_env_t h);

namespace nfsp {

int set_error(int code, const char *fmt, ...);

#define NFSP_CHECK_ARG(cond, ...)                                    \
    do {                                                             \
        if (!(cond)) return nfsp::set_error(NFSP_E_ARG, __VA_ARGS__); \
    } while (0)

#define NFSP_CUDA(call)                                                                                   \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return nfsp::set_error(NFSP_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                                   __FILE__, __LINE__);                                                   \
    } while (0)

#define NFSP_LAUNCH_CHECK() NFSP_CUDA(cudaGetLastError())

// floor(x * 2^32) clamped: the integer threshold a Philox u32 is compared with for P(event) = x
inline uint32_t frac_u32(double x) {
    if (!(x > 0.0)) return 0u;
    const double v = x * 4294967296.0;
    return v >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)v;
}

// grid for a grid-stride loop over n items: whole waves of the SM count, capped by the work
inline int grid_for(int64_t n, int threads, int sm_count, int ctas_per_sm) {
    const int64_t need = (n + threads - 1) / threads;
    const int64_t full = (int64_t)sm_count * ctas_per_sm;
    int64_t g = need < full ? need : full;
    if (g < 1) g = 1;
    return (int)g;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace nfsp
