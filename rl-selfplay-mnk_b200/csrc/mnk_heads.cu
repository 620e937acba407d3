// mnk_heads.cu -- the policy / value head tails of the reference's ResNet as ONE kernel.
//
// After the tcgen05 tower (mnk_resnet.cu) has produced Flatten(Conv2d(32,2,1)) -> policy_feat [N][2A]
// and Flatten(Conv2d(32,1,1)) -> value_feat [N][A], the heads continue (reference:
// src/alg/architectures/resnet.py:41-63):
//     policy: LayerNorm(2A) -> ReLU -> Linear(2A,128) -> LayerNorm(128) -> ReLU -> Linear(128,A)   = logits
//     value : LayerNorm(A)  -> ReLU -> Linear(A,128)  -> LayerNorm(128) -> ReLU -> Linear(128,1) -> Tanh
// 2% of the forward's FLOPs, but 14 small library kernels per forward when run through torch modules
// (they cost as much wall time as the whole tower at 32k envs).  Here a CTA of 256 threads takes 8
// samples at a time: warp w normalises sample w (shuffle reductions), the two hidden layers are
// computed with one thread per (hidden unit, 4 samples) reading the transposed fp32 weight matrices
// through the read-only path (coalesced over units, L1/L2 resident: 166 KB at 9x9) and the
// activations as float4 broadcasts from shared memory.  fp32 throughout (matches torch to ~1e-6).
#include "mnk_dispatch.cuh"

namespace hd {
constexpr int kH = 128;          // head_hidden_dim of resnet_b_s
constexpr int kSB = 8;           // samples per CTA iteration (= warps)
constexpr int kThreads = 256;
constexpr float kEps = 1e-5f;    // torch.nn.LayerNorm default

MNK_DEV float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(MNK_FULL_WARP, v, o);
    return v;
}

// LayerNorm + ReLU of `len` values of sample s held at x[k*kSB + s] (k-major so that the next layer can
// fetch 4 samples of one k with a single float4), in place; one warp per sample
MNK_DEV void layernorm_relu(float* x, int s, int len, const float* __restrict__ gamma, const float* __restrict__ beta, int lane) {
    float sum = 0.f;
    for (int k = lane; k < len; k += 32) sum += x[k * kSB + s];
    const float mean = warp_sum(sum) / (float)len;
    float sq = 0.f;
    for (int k = lane; k < len; k += 32) {
        const float d = x[k * kSB + s] - mean;
        sq += d * d;
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)len + kEps);
    for (int k = lane; k < len; k += 32) {
        const float y = (x[k * kSB + s] - mean) * rstd * __ldg(gamma + k) + __ldg(beta + k);
        x[k * kSB + s] = fmaxf(y, 0.f);
    }
}

// out[j][4 samples] = bias[j] + sum_k wT[k][j] * x[k][samples]   for j = unit, samples = 4*half .. 4*half+3
MNK_DEV void dense4(const float* __restrict__ wT, int ld, const float* __restrict__ bias, const float* x, int len, int unit,
                    int half, float (&acc)[4]) {
    const float b = __ldg(bias + unit);
    acc[0] = acc[1] = acc[2] = acc[3] = b;
    const float4* xv = reinterpret_cast<const float4*>(x) + half;      // x[k*8 + 4*half ..]
#pragma unroll 4
    for (int k = 0; k < len; ++k) {
        const float w = __ldg(wT + (size_t)k * ld + unit);
        const float4 v = xv[2 * k];
        acc[0] = fmaf(w, v.x, acc[0]);
        acc[1] = fmaf(w, v.y, acc[1]);
        acc[2] = fmaf(w, v.z, acc[2]);
        acc[3] = fmaf(w, v.w, acc[3]);
    }
}

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

    for (long long r0 = (long long)blockIdx.x * kSB; r0 < rows; r0 += (long long)gridDim.x * kSB) {
        // 1. load (coalesced per sample row) + LayerNorm + ReLU; rows past the end are zero-filled
        {
            const long long r = r0 + warp;
            const bool live = r < rows;
            for (int k = lane; k < two; k += 32) xp[k * kSB + warp] = live ? __ldg(pf + (size_t)r * two + k) : 0.f;
            for (int k = lane; k < cells; k += 32) xv[k * kSB + warp] = live ? __ldg(vf + (size_t)r * cells + k) : 0.f;
            __syncwarp();
            layernorm_relu(xp, warp, two, w.p_ln1_w, w.p_ln1_b, lane);
            layernorm_relu(xv, warp, cells, w.v_ln1_w, w.v_ln1_b, lane);
        }
        __syncthreads();
        // 2. first Linear of both heads: thread = (hidden unit, 4 samples)
        {
            float acc[4];
            dense4(w.p_w1t, kH, w.p_b1, xp, two, unit, half, acc);
            *reinterpret_cast<float4*>(hp + unit * kSB + 4 * half) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            dense4(w.v_w1t, kH, w.v_b1, xv, cells, unit, half, acc);
            *reinterpret_cast<float4*>(hv + unit * kSB + 4 * half) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
        __syncthreads();
        // 3. LayerNorm(128) + ReLU per sample
        layernorm_relu(hp, warp, kH, w.p_ln2_w, w.p_ln2_b, lane);
        layernorm_relu(hv, warp, kH, w.v_ln2_w, w.v_ln2_b, lane);
        __syncthreads();
        // 4. output layers: logits (thread = (cell, 4 samples)), value (warp = sample)
        for (int a = unit; a < cells; a += kH) {
            float acc[4];
            dense4(w.p_w2t, cells, w.p_b2, hp, kH, a, half, acc);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long r = r0 + 4 * half + i;
                if (r < rows) logits[(size_t)r * cells + a] = acc[i];
            }
        }
        {
            float dot = 0.f;
            for (int k = lane; k < kH; k += 32) dot = fmaf(hv[k * kSB + warp], __ldg(w.v_w2 + k), dot);
            dot = warp_sum(dot);
            const long long r = r0 + warp;
            if (lane == 0 && r < rows) values[r] = tanhf(dot + __ldg(w.v_b2));
        }
        __syncthreads();   // smem is rewritten by the next batch
    }
}
}  // namespace hd

extern "C" int mnk_resnet_heads(const float* policy_feat, const float* value_feat, int64_t rows, int32_t cells,
                                const mnk_heads_weights_t* w, float* logits, float* values, void* stream) {
    if (!policy_feat || !value_feat || !w || !logits || !values) return MNK_ERR_NULL;
    const float* const* ptrs = reinterpret_cast<const float* const*>(w);
    for (size_t i = 0; i < sizeof(mnk_heads_weights_t) / sizeof(float*); ++i)
        if (ptrs[i] == nullptr) return MNK_ERR_NULL;
    if (rows < 0 || cells < 1 || cells > 1024) return MNK_ERR_ARG;
    if (rows == 0) return MNK_OK;
    const size_t smem = sizeof(float) * hd::kSB * (size_t)(3 * cells + 2 * hd::kH);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(hd::heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    const long long batches = (rows + hd::kSB - 1) / hd::kSB;
    const unsigned grid = (unsigned)(batches < 148 * 4 ? batches : 148 * 4);
    hd::heads_kernel<<<grid, hd::kThreads, smem, static_cast<cudaStream_t>(stream)>>>(policy_feat, value_feat, rows, cells, *w,
                                                                                     logits, values);
    return mnk_launch_status();
}
