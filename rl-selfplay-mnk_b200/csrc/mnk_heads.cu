// mnk_heads.cu -- the policy / value head tails of the reference's ResNet as ONE kernel.
//
// After the tcgen05 tower (mnk_resnet.cu) has produced Flatten(Conv2d(32,2,1)) -> policy_feat [N][2A]
// and Flatten(Conv2d(32,1,1)) -> value_feat [N][A], the heads continue (reference:
// src/alg/architectures/resnet.py:41-63):
//     policy: LayerNorm(2A) -> ReLU -> Linear(2A,128) -> LayerNorm(128) -> ReLU -> Linear(128,A)   = logits
//     value : LayerNorm(A)  -> ReLU -> Linear(A,128)  -> LayerNorm(128) -> ReLU -> Linear(128,1) -> Tanh
// 2% of the forward's FLOPs, but 14 small library kernels per forward when run through torch modules
// (they cost as much wall time as the whole tower at 32k envs).  Here a CTA of 256 threads takes 16
// samples at a time: warp w normalises samples w and w+8 (shuffle reductions), the two hidden layers are
// computed with one thread per (hidden unit, 8 samples) reading the transposed fp32 weight matrices
// through the read-only path (coalesced over units, L1/L2 resident: 166 KB at 9x9) and the
// activations as float4 broadcasts from shared memory (k-major, 16 samples per k).  fp32 throughout (matches torch to ~1e-6).
#include "mnk_dispatch.cuh"

namespace hd {
constexpr int kH = 128;          // head_hidden_dim of resnet_b_s
constexpr int kSB = 16;          // samples per CTA iteration (2 per warp; 8 per thread in the dense layers)
constexpr int kThreads = 256;
constexpr float kEps = 1e-5f;    // torch.nn.LayerNorm default

MNK_DEV float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(MNK_FULL_WARP, v, o);
    return v;
}

// LayerNorm + ReLU of `len` values of sample s held at x[k*kSB + s] (k-major so that the next layer can
// fetch 8 samples of one k with two float4), in place; one warp per sample.  The values are pulled into
// registers once (the k-major layout makes lane-strided shared-memory passes 16-way bank conflicted).
constexpr int kMaxItems = 24;    // per lane: covers len <= 768 (19x19: 2A = 722)

template <int kItems>
MNK_DEV void layernorm_relu_regs(float (&v)[kItems], int len, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float* x, int s, int lane) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kItems; ++i) sum += (lane + 32 * i < len) ? v[i] : 0.f;
    const float mean = warp_sum(sum) / (float)len;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const float d = (lane + 32 * i < len) ? v[i] - mean : 0.f;
        sq += d * d;
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)len + kEps);
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int k = lane + 32 * i;
        if (k < len) x[k * kSB + s] = fmaxf((v[i] - mean) * rstd * __ldg(gamma + k) + __ldg(beta + k), 0.f);
    }
}

// source = global row (coalesced) or the k-major shared-memory column of sample s.  kItems = register
// slots per lane the caller guarantees to be enough (len <= 32 * kItems) -- 0 = unknown, shared-memory passes.
template <int kItems, bool kFromGlobal>
MNK_DEV void layernorm_relu(const float* __restrict__ grow, bool live, float* x, int s, int len,
                            const float* __restrict__ gamma, const float* __restrict__ beta, int lane) {
    if constexpr (kItems > 0) {
        float v[kItems];
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const int k = lane + 32 * i;
            if constexpr (kFromGlobal)
                v[i] = (k < len && live) ? __ldg(grow + k) : 0.f;
            else
                v[i] = (k < len) ? x[k * kSB + s] : 0.f;
        }
        layernorm_relu_regs<kItems>(v, len, gamma, beta, x, s, lane);
        return;
    }
    // very large boards: shared-memory passes
    if constexpr (kFromGlobal)
        for (int k = lane; k < len; k += 32) x[k * kSB + s] = live ? __ldg(grow + k) : 0.f;
    __syncwarp();
    float sum = 0.f;
    for (int k = lane; k < len; k += 32) sum += x[k * kSB + s];
    const float mean = warp_sum(sum) / (float)len;
    float sq = 0.f;
    for (int k = lane; k < len; k += 32) {
        const float d = x[k * kSB + s] - mean;
        sq += d * d;
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)len + kEps);
    for (int k = lane; k < len; k += 32)
        x[k * kSB + s] = fmaxf((x[k * kSB + s] - mean) * rstd * __ldg(gamma + k) + __ldg(beta + k), 0.f);
}

// out[j][8 samples] = bias[j] + sum_k wT[k][j] * x[k][samples]   for j = unit, samples = 8*half .. 8*half+7
// The weight column is walked in groups of four k with the NEXT group's weights already in flight while the
// current 32 FMAs issue (the loads are L1 / L2 hits whose latency was the kernel's top stall); two register
// sets alternate so that nothing is moved, the steady-state loop carries no predicates or index clamps
// (47 instructions per 32 FMAs), and the last groups / the len % 4 tail are peeled.
MNK_DEV void fma_group(const float (&w)[4], const float4* __restrict__ xv, float (&acc)[8]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 v0 = xv[4 * j], v1 = xv[4 * j + 1];
        acc[0] = fmaf(w[j], v0.x, acc[0]);
        acc[1] = fmaf(w[j], v0.y, acc[1]);
        acc[2] = fmaf(w[j], v0.z, acc[2]);
        acc[3] = fmaf(w[j], v0.w, acc[3]);
        acc[4] = fmaf(w[j], v1.x, acc[4]);
        acc[5] = fmaf(w[j], v1.y, acc[5]);
        acc[6] = fmaf(w[j], v1.z, acc[6]);
        acc[7] = fmaf(w[j], v1.w, acc[7]);
    }
}

MNK_DEV void load_group(float (&w)[4], const float* __restrict__ wp, int ld) {
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = __ldg(wp + (size_t)j * ld);
}

MNK_DEV void dense8(const float* __restrict__ wT, int ld, const float* __restrict__ bias, const float* x, int len, int unit,
                    int half, float (&acc)[8]) {
    const float b = __ldg(bias + unit);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = b;
    const float4* xv = reinterpret_cast<const float4*>(x) + 2 * half;      // x[k*16 + 8*half ..]: 4 float4 per k
    const float* wp = wT + unit;
    const size_t step = (size_t)4 * ld;
    int groups = len >> 2;
    if (groups > 0) {
        float wa[4], wb[4];
        load_group(wa, wp, ld);
        while (groups >= 3) {              // groups g, g+1 computed; g+1, g+2 fetched
            load_group(wb, wp + step, ld);
            fma_group(wa, xv, acc);
            load_group(wa, wp + 2 * step, ld);
            fma_group(wb, xv + 16, acc);
            wp += 2 * step;
            xv += 32;
            groups -= 2;
        }
        if (groups == 2) {
            load_group(wb, wp + step, ld);
            fma_group(wa, xv, acc);
            fma_group(wb, xv + 16, acc);
        } else {
            fma_group(wa, xv, acc);
        }
        wp += (size_t)groups * step;
        xv += 16 * groups;
    }
    for (int k = len & ~3; k < len; ++k, wp += ld, xv += 4) {
        const float w = __ldg(wp);
        const float4 v0 = xv[0], v1 = xv[1];
        acc[0] = fmaf(w, v0.x, acc[0]);
        acc[1] = fmaf(w, v0.y, acc[1]);
        acc[2] = fmaf(w, v0.z, acc[2]);
        acc[3] = fmaf(w, v0.w, acc[3]);
        acc[4] = fmaf(w, v1.x, acc[4]);
        acc[5] = fmaf(w, v1.y, acc[5]);
        acc[6] = fmaf(w, v1.z, acc[6]);
        acc[7] = fmaf(w, v1.w, acc[7]);
    }
}

// kItems: register slots per lane for the policy features (2A <= 32 * kItems); the value features use half.
template <int kItems>
__global__ void __launch_bounds__(kThreads)
heads_kernel(const float* __restrict__ pf, const float* __restrict__ vf, long long rows, int cells, mnk_heads_weights_t w,
             float* __restrict__ logits, float* __restrict__ values) {
    extern __shared__ __align__(16) float smem[];
    const int two = 2 * cells;
    float* xp = smem;                       // [2A][8]  policy features -> LN1 output
    float* xv = xp + (size_t)two * kSB;     // [A][8]   value features  -> LN1 output
    float* hp = xv + (size_t)cells * kSB;   // [128][8] policy hidden
    float* hv = hp + kH * kSB;              // [128][8] value hidden
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int unit = tid & (kH - 1), half = tid >> 7;
    const bool want_value = values != nullptr;       // policy-only callers (the opponent) skip the value head

    for (long long r0 = (long long)blockIdx.x * kSB; r0 < rows; r0 += (long long)gridDim.x * kSB) {
        // 1. load (coalesced per sample row) + LayerNorm + ReLU; rows past the end are zero-filled
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
            const int sidx = warp + 8 * rep;
            const long long r = r0 + sidx;
            const bool live = r < rows;
            const long long rr = live ? r : 0;
            layernorm_relu<kItems, true>(pf + (size_t)rr * two, live, xp, sidx, two, w.p_ln1_w, w.p_ln1_b, lane);
            if (want_value)
                layernorm_relu<(kItems + 1) / 2, true>(vf + (size_t)rr * cells, live, xv, sidx, cells, w.v_ln1_w, w.v_ln1_b, lane);
        }
        __syncthreads();
        // 2. first Linear of both heads: thread = (hidden unit, 8 samples)
        {
            float acc[8];
            dense8(w.p_w1t, kH, w.p_b1, xp, two, unit, half, acc);
            float4* dst = reinterpret_cast<float4*>(hp + unit * kSB + 8 * half);
            dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            if (want_value) {
                dense8(w.v_w1t, kH, w.v_b1, xv, cells, unit, half, acc);
                dst = reinterpret_cast<float4*>(hv + unit * kSB + 8 * half);
                dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            }
        }
        __syncthreads();
        // 3. LayerNorm(128) + ReLU per sample
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
            layernorm_relu<kH / 32, false>(nullptr, true, hp, warp + 8 * rep, kH, w.p_ln2_w, w.p_ln2_b, lane);
            if (want_value) layernorm_relu<kH / 32, false>(nullptr, true, hv, warp + 8 * rep, kH, w.v_ln2_w, w.v_ln2_b, lane);
        }
        __syncthreads();
        // 4. output layers: logits (thread = (cell, 8 samples)), value (warp = 2 samples)
        for (int a = unit; a < cells; a += kH) {
            float acc[8];
            dense8(w.p_w2t, cells, w.p_b2, hp, kH, a, half, acc);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long r = r0 + 8 * half + i;
                if (r < rows) logits[(size_t)r * cells + a] = acc[i];
            }
        }
#pragma unroll
        for (int rep = 0; rep < 2 && want_value; ++rep) {
            const int sidx = warp + 8 * rep;
            float dot = 0.f;
            for (int k = lane; k < kH; k += 32) dot = fmaf(hv[k * kSB + sidx], __ldg(w.v_w2 + k), dot);
            dot = warp_sum(dot);
            const long long r = r0 + sidx;
            if (lane == 0 && r < rows) values[r] = tanhf(dot + __ldg(w.v_b2));
        }
        __syncthreads();   // smem is rewritten by the next batch
    }
}
}  // namespace hd

extern "C" int mnk_resnet_heads(const float* policy_feat, const float* value_feat, int64_t rows, int32_t cells,
                                const mnk_heads_weights_t* w, float* logits, float* values, void* stream) {
    if (!policy_feat || !w || !logits || (values && !value_feat)) return MNK_ERR_NULL;
    const float* const* ptrs = reinterpret_cast<const float* const*>(w);
    for (size_t i = 0; i < sizeof(mnk_heads_weights_t) / sizeof(float*); ++i)
        if (ptrs[i] == nullptr) return MNK_ERR_NULL;
    if (rows < 0 || cells < 1 || cells > 1024) return MNK_ERR_ARG;
    if (rows == 0) return MNK_OK;
    const size_t smem = sizeof(float) * hd::kSB * (size_t)(3 * cells + 2 * hd::kH);
    const long long batches = (rows + hd::kSB - 1) / hd::kSB;
    auto launch = [&](auto kernel, std::atomic<size_t>* granted) {
        if (int rc = mnk_optin_smem(kernel, smem, granted)) return rc;
        // persistent grid: exactly the CTAs that are resident at once on the device's SMs (one wave, no tail)
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, hd::kThreads, smem);
        if (e != cudaSuccess) return (int)e;
        const long long resident = (long long)mnk_sm_count() * (per_sm > 0 ? per_sm : 1);
        const unsigned grid = (unsigned)(batches < resident ? batches : resident);
        kernel<<<grid, hd::kThreads, smem, static_cast<cudaStream_t>(stream)>>>(policy_feat, value_feat, rows, cells, *w, logits,
                                                                                values);
        return mnk_launch_status();
    };
    static std::atomic<size_t> granted[5][kMaxDevices];
    const int items = (2 * cells + 31) / 32;       // LayerNorm register slots per lane
    if (items <= 6) return launch(hd::heads_kernel<6>, granted[0]);      // up to 9x9 (2A = 162)
    if (items <= 11) return launch(hd::heads_kernel<11>, granted[1]);    // 13x13 (338)
    if (items <= 15) return launch(hd::heads_kernel<15>, granted[2]);    // 15x15 (450)
    if (items <= hd::kMaxItems) return launch(hd::heads_kernel<hd::kMaxItems>, granted[3]);   // 19x19 (722)
    return launch(hd::heads_kernel<0>, granted[4]);                      // larger: shared-memory passes
}
