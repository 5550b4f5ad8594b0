// qt_policy.cu — the deep tracking policy evaluated inside the closed loop.
//
// Replaces, for --controller deep and for imitation training,
//     features = controller._extract_features(obs); action = network(features); env.step(action)
// (controllers/deep_tracking_policy.py:199-286, eval.py:119-165, train.py:668-721) by one kernel
// launch for all episodes: one thread per episode, the env in fp64 exactly as in qt_rollout.cu (same
// device functions), the network in fp32 like torch (the parity target is torch's fp32 Linear, so it
// stays on the FP32 pipe: TF32 / BF16 tensor-core products would miss that target by 1e-3).
//
// Two evaluations of the network, identical bit for bit (same products added in the same order):
//  * TILED — the reference's default 64-64 network and every shape whose hidden widths are within 25 %
//    of multiples of 64.  The 32 episodes of a warp x the units of a layer form a small GEMM; every
//    lane owns an 8-episode x 8-unit register tile, one k-step is four 16-byte shared-memory loads for
//    32 packed FFMA2 (= 64 FMAs); activations travel between layers through per-warp shared-memory
//    buffers with an XOR swizzle that keeps every access pattern conflict free (tile_layer,
//    mlp_forward_tiled below).  64-64: 22 KB of staged weights + 8 KB of activations per warp, three
//    resident blocks per SM.
//  * PER-THREAD — any other shape and qt_policy_action_batched.  The packed weights are staged in
//    shared memory and read by every thread at the same address (broadcast loads); the 18 features
//    stay in registers for the first layer; hidden activations live in the thread's own column of a
//    shared-memory tile; every chunk of 16 output units accumulates in registers (one weight vector
//    load per two FFMA2); the last hidden layer is never stored — each chunk of it is multiplied
//    into the four output accumulators as soon as it is activated.
#include "qt_common.cuh"
#include "qt_policy_tc.cuh"
#include "qt_step.cuh"

namespace qt {

constexpr int kPolicyThreads = 128;
#ifndef QT_POLICY_TILED_3
#define QT_POLICY_TILED_3 1
#endif

__host__ __device__ inline int pad4(int x) { return (x + 3) & ~3; }

struct MlpDesc {
    int n_layers;
    int rows[QT_MLP_MAX_LAYERS];     // input rows of layer l in the packed layout
    int outp[QT_MLP_MAX_LAYERS];     // padded output width of layer l
    int w_off[QT_MLP_MAX_LAYERS];    // offset (floats) of Wt of layer l; the bias follows Wt
    int activation;
    int total_floats;
    int tile_w[2];                   // widths (floats per thread) of the two activation tiles: hidden layer l
                                     // (0-based) is stored in tile l & 1; the last hidden layer is not stored
    float out_lo[4], out_scale[4];
    // TILED evaluation: hidden widths padded to multiples of 64 (zero weights and biases, so a padded unit
    // outputs activation(0) = 0 and feeds zero rows of the next layer)
    int tiled;                       // 1 if the tiled evaluation applies to this shape
    int t_rows[QT_MLP_MAX_LAYERS];   // input rows of layer l in the staged (tiled) layout
    int t_n[QT_MLP_MAX_LAYERS];      // output width of layer l in the staged layout (last layer: 4)
    int t_off[QT_MLP_MAX_LAYERS];    // offset (floats) of Wt of layer l in the staged layout; bias follows
    int t_total;                     // floats of the staged weights
    int t_buf[2];                    // rows (of 32 floats) of the two per-warp activation buffers
    int t_inplace;                   // every hidden layer is one 64-unit column tile: the layers run in place in buffer 0
    int tc_ok;                       // one to three hidden layers of at most 64 units: the tensor-core evaluation applies
    tc::TcLayout tc;                 // its shared-memory plan (qt_policy_tc.cuh)
};

static bool describe(const qt_mlp &m, MlpDesc &d) {
    if (m.n_layers < 1 || m.n_layers > QT_MLP_MAX_LAYERS) return false;
    if (m.dims[0] != QT_FEATURE_DIM || m.dims[m.n_layers] != QT_ACTION_DIM) return false;
    if (m.activation < QT_ACT_RELU || m.activation > QT_ACT_LEAKY_RELU) return false;
    d.n_layers = m.n_layers;
    d.activation = m.activation;
    int off = 0;
    d.tile_w[0] = d.tile_w[1] = 0;
    for (int l = 0; l < m.n_layers; l++) {
        if (m.dims[l + 1] < 1 || m.dims[l + 1] > QT_MLP_MAX_WIDTH) return false;
        d.rows[l] = l == 0 ? m.dims[0] : pad4(m.dims[l]);
        d.outp[l] = pad4(m.dims[l + 1]);
        d.w_off[l] = off;
        off += d.rows[l] * d.outp[l] + d.outp[l];
        if (l < m.n_layers - 2 && d.outp[l] > d.tile_w[l & 1]) d.tile_w[l & 1] = d.outp[l];  // stored hidden layers
    }
    d.total_floats = off;
    for (int i = 0; i < 4; i++) {
        d.out_lo[i] = m.out_lo[i];
        d.out_scale[i] = m.out_scale[i];
    }
    // tiled layout: layer l reads activation buffer l & 1 and writes buffer (l + 1) & 1; the features are in buffer 0
    int toff = 0;
    long fma_tiled = 0, fma_packed = 0;
    d.t_buf[0] = QT_FEATURE_DIM;
    d.t_buf[1] = 0;
    for (int l = 0; l < m.n_layers; l++) {
        const bool last = l == m.n_layers - 1;
        d.t_rows[l] = l == 0 ? QT_FEATURE_DIM : d.t_n[l - 1];
        d.t_n[l] = last ? 4 : ((m.dims[l + 1] + 63) & ~63);
        d.t_off[l] = toff;
        toff += d.t_rows[l] * d.t_n[l] + d.t_n[l];
        if (!last && d.t_n[l] > d.t_buf[(l + 1) & 1]) d.t_buf[(l + 1) & 1] = d.t_n[l];
        fma_tiled += (long)d.t_rows[l] * d.t_n[l];
        fma_packed += (long)d.rows[l] * d.outp[l];
    }
    d.t_total = (toff + 3) & ~3;
    d.t_inplace = 1;
    for (int l = 0; l + 1 < m.n_layers; l++) d.t_inplace &= d.t_n[l] == 64;
    if (d.t_inplace) {   // a layer's outputs are complete in registers before the first of them is stored
        d.t_buf[0] = 64;
        d.t_buf[1] = 0;
    }
    d.t_buf[0] += 2;     // tile_layer's operand prefetch runs two rows past the layer's input
    if (d.t_buf[1]) d.t_buf[1] += 2;
    d.tiled = m.n_layers >= 2 && fma_tiled * 4 <= fma_packed * 5;   // at most 25 % padding work
    d.tc_ok = m.n_layers >= 2 && m.n_layers - 1 <= tc::kMaxHidden;
    for (int l = 1; l < m.n_layers; l++) d.tc_ok = d.tc_ok && m.dims[l] <= tc::kN;
    if (d.tc_ok) {
        tc::TcLayout &t = d.tc;
        t.hidden = m.n_layers - 1;
        int at = tc::kWeights;
        for (int l = 0; l < t.hidden; l++) {
            t.n[l] = (m.dims[l + 1] + 15) & ~15;
            t.k[l] = l == 0 ? tc::kK0 : t.n[l - 1];
            t.b_hi[l] = at;
            t.b_lo[l] = at + t.k[l] * t.n[l];
            at += 2 * t.k[l] * t.n[l];
        }
        t.w3 = at;
        t.bias = t.w3 + tc::kN * 4;
        t.bar = t.bias + tc::kMaxHidden * tc::kN + 4;
        t.bar += t.bar & 1;                      // 8-byte alignment of the mbarriers
        t.tmem_slot = t.bar + 2 * tc::kSets;
        t.floats = t.tmem_slot + 2;
    }
    return true;
}

// Two fp32 FMAs in one instruction (sm_100 FFMA2): acc.lo += a * w.lo, acc.hi += a * w.hi, each rounded
// exactly like fmaf.  The FP32 pipe does the same work either way; what halves is the number of issue
// slots the 5 504 FMAs of a 64-64 network take, and the tiled kernel is issue-bound.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(float a, unsigned long long w, unsigned long long &acc) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(pack2(a, a)), "l"(w));   // in place: no accumulator copies
}

template <int CH>
__device__ __forceinline__ void chunk_bias(const float *__restrict__ bias, int j0, float (&acc)[CH]) {
#pragma unroll
    for (int i = 0; i < CH; i += 4) {
        const float4 b = *reinterpret_cast<const float4 *>(bias + j0 + i);
        acc[i] = b.x;
        acc[i + 1] = b.y;
        acc[i + 2] = b.z;
        acc[i + 3] = b.w;
    }
}

template <int CH>
__device__ __forceinline__ void chunk_fma(const float *__restrict__ w, float xk, float (&acc)[CH]) {
#pragma unroll
    for (int i = 0; i < CH; i += 4) {
        const float4 wk = *reinterpret_cast<const float4 *>(w + i);
        unsigned long long lo = pack2(acc[i], acc[i + 1]), hi = pack2(acc[i + 2], acc[i + 3]);
        ffma2(xk, pack2(wk.x, wk.y), lo);
        ffma2(xk, pack2(wk.z, wk.w), hi);
        unpack2(lo, acc[i], acc[i + 1]);
        unpack2(hi, acc[i + 2], acc[i + 3]);
    }
}

// One chunk of CH (16 or 4) output units of one layer for this thread; the layer input is this thread's
// column of a shared-memory tile (xin[k * T]).
template <int CH, int T>
__device__ __forceinline__ void mlp_chunk(const float *__restrict__ Wt, const float *__restrict__ bias, int rows,
                                          int outp, int j0, const float *xin, float (&acc)[CH]) {
    chunk_bias<CH>(bias, j0, acc);
    const float *w = Wt + j0;
#pragma unroll 4
    for (int k = 0; k < rows; k++) {
        chunk_fma<CH>(w, xin[k * T], acc);
        w += outp;
    }
}

// The same for the first layer: its 18 inputs are in registers.
template <int CH>
__device__ __forceinline__ void mlp_chunk_first(const float *__restrict__ Wt, const float *__restrict__ bias,
                                                int outp, int j0, const float (&feat)[QT_FEATURE_DIM], float (&acc)[CH]) {
    chunk_bias<CH>(bias, j0, acc);
    const float *w = Wt + j0;
#pragma unroll
    for (int k = 0; k < QT_FEATURE_DIM; k++) chunk_fma<CH>(w + k * outp, feat[k], acc);
}

template <int KIND>
__device__ __forceinline__ float activate_t(float x) {
    if (KIND == QT_ACT_TANH) return tanhf(x);
    if (KIND == QT_ACT_ELU) return x > 0.0f ? x : expm1f(x);
    if (KIND == QT_ACT_LEAKY_RELU) return x > 0.0f ? x : 0.1f * x;
    float r;   // torch.relu propagates NaN, fmaxf does not; max.NaN does, in one instruction
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(x));
    return r;
}

// Activate a finished chunk and either store it (hidden layer) or multiply it into the output
// accumulators (last hidden layer, fused output layer).  The activation kind is dispatched once per
// chunk, not per element.
template <int KIND, int CH, int T>
__device__ __forceinline__ void finish_chunk_t(const float (&acc)[CH], bool last_hidden, const float *__restrict__ Wo,
                                               int j0, float *xout, float (&o)[4]) {
    if (last_hidden) {
#pragma unroll
        for (int i = 0; i < CH; i++) chunk_fma<4>(Wo + (j0 + i) * 4, activate_t<KIND>(acc[i]), o);
    } else {
#pragma unroll
        for (int i = 0; i < CH; i++) xout[(j0 + i) * T] = activate_t<KIND>(acc[i]);
    }
}

template <int CH, int T>
__device__ __forceinline__ void finish_chunk(int kind, const float (&acc)[CH], bool last_hidden,
                                             const float *__restrict__ Wo, int j0, float *xout, float (&o)[4]) {
    switch (kind) {
    case QT_ACT_TANH: finish_chunk_t<QT_ACT_TANH, CH, T>(acc, last_hidden, Wo, j0, xout, o); break;
    case QT_ACT_ELU: finish_chunk_t<QT_ACT_ELU, CH, T>(acc, last_hidden, Wo, j0, xout, o); break;
    case QT_ACT_LEAKY_RELU: finish_chunk_t<QT_ACT_LEAKY_RELU, CH, T>(acc, last_hidden, Wo, j0, xout, o); break;
    default: finish_chunk_t<QT_ACT_RELU, CH, T>(acc, last_hidden, Wo, j0, xout, o); break;
    }
}

// PolicyNetwork.forward (deep_tracking_policy.py:114-147) for one feature vector held in registers.
// tile0 / tile1: this thread's columns of the two activation tiles.  W: packed weights (shared or global).
template <int T>
__device__ __forceinline__ void mlp_forward(const MlpDesc &M, const float *__restrict__ W, float *tile0, float *tile1,
                                            const float (&feat)[QT_FEATURE_DIM], float (&out)[4]) {
    const int L = M.n_layers;                      // linear layers; hidden layers are 0 .. L-2
    const float *Wo = W + M.w_off[L - 1];          // output layer: Wt [rows][4], bias [4]
    const int rows_o = M.rows[L - 1];
    float o[4];
    chunk_bias<4>(Wo + rows_o * 4, 0, o);
    if (L == 1) {                                  // no hidden layer: a linear policy
#pragma unroll
        for (int k = 0; k < QT_FEATURE_DIM; k++) chunk_fma<4>(Wo + k * 4, feat[k], o);
    }
    for (int l = 0; l + 1 < L; l++) {
        const float *Wt = W + M.w_off[l];
        const int rows = M.rows[l], outp = M.outp[l];
        const float *bias = Wt + rows * outp;
        const float *xin = ((l - 1) & 1) ? tile1 : tile0;
        float *xout = (l & 1) ? tile1 : tile0;
        const bool last_hidden = (l == L - 2);
        int j0 = 0;
        for (; j0 + 16 <= outp; j0 += 16) {
            float acc[16];
            if (l == 0)
                mlp_chunk_first<16>(Wt, bias, outp, j0, feat, acc);
            else
                mlp_chunk<16, T>(Wt, bias, rows, outp, j0, xin, acc);
            finish_chunk<16, T>(M.activation, acc, last_hidden, Wo, j0, xout, o);
        }
        for (; j0 < outp; j0 += 4) {
            float acc[4];
            if (l == 0)
                mlp_chunk_first<4>(Wt, bias, outp, j0, feat, acc);
            else
                mlp_chunk<4, T>(Wt, bias, rows, outp, j0, xin, acc);
            finish_chunk<4, T>(M.activation, acc, last_hidden, Wo, j0, xout, o);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float sg = 1.0f / (1.0f + expf(-o[i]));  // torch.sigmoid
        out[i] = fmaf(sg, M.out_scale[i], M.out_lo[i]);
    }
}

// ---- TILED evaluation ----------------------------------------------------------------------------
// A warp's 32 episodes x the N units of a layer is a small GEMM.  Lane (er, jc) = (lane / 8, lane % 8)
// owns episodes {4 er .. 4 er + 3, 16 + 4 er .. 16 + 4 er + 3} and, of every 64-unit column tile, units
// {4 jc .. 4 jc + 3, 32 + 4 jc .. 32 + 4 jc + 3}: 64 accumulators.  One k-step loads two float4 of
// activations and two float4 of weights and issues 64 FFMAs (the per-thread evaluation needs one
// float4 of weights per 4 FFMAs and is bound by the shared-memory pipe and its latency).  Every
// accumulator starts from its bias and adds x_k * w_kj in ascending k, exactly as the per-thread
// evaluation does, so the two give identical bits.
// Activation buffers are per warp, [unit][episode] with 32 floats per unit; the 16-byte chunk c of
// row j is stored at chunk c ^ ((j >> 2) & 7), which makes the transposed float4 stores of the
// epilogue, the float4 loads of the k-loop and the per-episode scalar accesses all conflict free.
__device__ __forceinline__ int tile_elem(int row, int lane) {   // float index of (unit row, episode lane)
    return row * 32 + ((((lane >> 2) ^ (row >> 2)) & 7) << 2) + (lane & 3);
}

template <int KIND>
__device__ __forceinline__ void tile_layer(const float *__restrict__ Wt, const float *__restrict__ bias, int K, int N,
                                           const float *xin, float *xout, int er, int jc) {
    const bool inplace = xin == xout;   // only with a single column tile (MlpDesc::t_inplace)
    for (int j0 = 0; j0 < N; j0 += 64) {
        unsigned long long acc[8][4];   // [episode][unit pair]
        {
            const float4 b0 = *reinterpret_cast<const float4 *>(bias + j0 + jc * 4);
            const float4 b1 = *reinterpret_cast<const float4 *>(bias + j0 + 32 + jc * 4);
            const unsigned long long b[4] = {pack2(b0.x, b0.y), pack2(b0.z, b0.w), pack2(b1.x, b1.y), pack2(b1.z, b1.w)};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[i][c] = b[c];
        }
        const float *wp = Wt + j0 + jc * 4;
        // Software pipeline: two operand sets (A, B), each reloaded in place right after its last use, so
        // the loads of step k + 2 fly during the 32 FFMA2 of step k + 1.  Rows are walked in groups of
        // four (one swizzle per group) plus a tail of two (K is 18 or a multiple of 64).
        struct Operands {
            float4 a0, a1, w0, w1;
        };
        auto load = [&](const float *xr, const float *wr, int o0, int o1, Operands &v) {
            v.a0 = *reinterpret_cast<const float4 *>(xr + o0);
            v.a1 = *reinterpret_cast<const float4 *>(xr + o1);
            v.w0 = *reinterpret_cast<const float4 *>(wr);
            v.w1 = *reinterpret_cast<const float4 *>(wr + 32);
        };
        auto fma_all = [&](const Operands &v) {
            const float a[8] = {v.a0.x, v.a0.y, v.a0.z, v.a0.w, v.a1.x, v.a1.y, v.a1.z, v.a1.w};
            const unsigned long long w[4] = {pack2(v.w0.x, v.w0.y), pack2(v.w0.z, v.w0.w), pack2(v.w1.x, v.w1.y),
                                             pack2(v.w1.z, v.w1.w)};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int c = 0; c < 4; c++) ffma2(a[i], w[c], acc[i][c]);
        };
        const int G = K >> 2;
        const bool tail = (K & 3) != 0;
        const float *xr = xin, *wr = wp;     // first row of the current group
        int o0 = er << 2, o1 = (4 + er) << 2;  // swizzled chunk offsets of the current group
        Operands A, B;
        load(xr, wr, o0, o1, A);
        load(xr + 32, wr + N, o0, o1, B);
#pragma unroll 1
        for (int g = 0; g < G; g++) {
            fma_all(A);
            load(xr + 64, wr + 2 * N, o0, o1, A);
            fma_all(B);
            load(xr + 96, wr + 3 * N, o0, o1, B);
            // after the last group these loads fetch rows K and K + 1, which exist (two spare rows per
            // activation buffer; the bias row and the next layer follow the weights) and are not used
            xr += 128;
            wr += 4 * N;
            const int sw = (g + 1) & 7;
            o0 = (er ^ sw) << 2;
            o1 = ((4 + er) ^ sw) << 2;
            fma_all(A);
            load(xr, wr, o0, o1, A);
            fma_all(B);
            load(xr + 32, wr + N, o0, o1, B);
        }
        if (tail) {
            fma_all(A);
            fma_all(B);
        }
        if (inplace) __syncwarp();   // every lane has read the inputs it needs
        const int sw = ((j0 >> 2) + jc) & 7;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float v[8][2];
#pragma unroll
            for (int i = 0; i < 8; i++) unpack2(acc[i][c], v[i][0], v[i][1]);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int j = j0 + (c < 2 ? jc * 4 + 2 * c + h : 32 + jc * 4 + 2 * (c - 2) + h);
                float *row = xout + j * 32;
                *reinterpret_cast<float4 *>(row + ((er ^ sw) << 2)) =
                    make_float4(activate_t<KIND>(v[0][h]), activate_t<KIND>(v[1][h]), activate_t<KIND>(v[2][h]),
                                activate_t<KIND>(v[3][h]));
                *reinterpret_cast<float4 *>(row + (((4 + er) ^ sw) << 2)) =
                    make_float4(activate_t<KIND>(v[4][h]), activate_t<KIND>(v[5][h]), activate_t<KIND>(v[6][h]),
                                activate_t<KIND>(v[7][h]));
            }
        }
    }
}

// PolicyNetwork.forward for the 32 feature vectors of a warp (one per lane); every lane of the warp must
// call it.  W: staged tiled weights (shared memory); buf0 / buf1: this warp's activation buffers.
__device__ __forceinline__ void mlp_forward_tiled(const MlpDesc &M, const float *__restrict__ W, float *buf0, float *buf1,
                                                  const float (&feat)[QT_FEATURE_DIM], float (&out)[4]) {
    const int lane = threadIdx.x & 31, er = lane >> 3, jc = lane & 7;
    const int L = M.n_layers;
    __syncwarp();   // the previous evaluation's output layer has finished reading the buffers
#pragma unroll
    for (int k = 0; k < QT_FEATURE_DIM; k++) buf0[tile_elem(k, lane)] = feat[k];
    __syncwarp();
    for (int l = 0; l + 1 < L; l++) {
        const float *Wt = W + M.t_off[l];
        const int K = M.t_rows[l], N = M.t_n[l];
        const float *xin = (l & 1) && !M.t_inplace ? buf1 : buf0;
        float *xout = (l & 1) || M.t_inplace ? buf0 : buf1;
        switch (M.activation) {
        case QT_ACT_TANH: tile_layer<QT_ACT_TANH>(Wt, Wt + K * N, K, N, xin, xout, er, jc); break;
        case QT_ACT_ELU: tile_layer<QT_ACT_ELU>(Wt, Wt + K * N, K, N, xin, xout, er, jc); break;
        case QT_ACT_LEAKY_RELU: tile_layer<QT_ACT_LEAKY_RELU>(Wt, Wt + K * N, K, N, xin, xout, er, jc); break;
        default: tile_layer<QT_ACT_RELU>(Wt, Wt + K * N, K, N, xin, xout, er, jc); break;
        }
        __syncwarp();
    }
    // output layer: four units, evaluated by each lane for its own episode
    const float *xin = ((L - 1) & 1) && !M.t_inplace ? buf1 : buf0;
    const float *Wo = W + M.t_off[L - 1];
    const int rows_o = M.t_rows[L - 1];
    float o[4];
    {
        const float4 b = *reinterpret_cast<const float4 *>(Wo + rows_o * 4);
        unsigned long long o01 = pack2(b.x, b.y), o23 = pack2(b.z, b.w);
        const float *xl = xin + (lane & 3);
        for (int g = 0; g < (rows_o >> 2); g++) {   // rows_o is a multiple of 64; one swizzle per four rows
            const float *xr = xl + g * 128 + ((((lane >> 2) ^ g) & 7) << 2);
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const float4 w = *reinterpret_cast<const float4 *>(Wo + (g * 4 + r) * 4);
                const float x = xr[r * 32];
                ffma2(x, pack2(w.x, w.y), o01);
                ffma2(x, pack2(w.z, w.w), o23);
            }
        }
        unpack2(o01, o[0], o[1]);
        unpack2(o23, o[2], o[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float sg = 1.0f / (1.0f + expf(-o[i]));  // torch.sigmoid
        out[i] = fmaf(sg, M.out_scale[i], M.out_lo[i]);
    }
}

// Stage the packed weights (global) into the tiled layout at the start of shared memory.
__device__ __forceinline__ void stage_weights_tiled(const MlpDesc &M, const float *__restrict__ Wg, float *smem_w) {
    for (int l = 0; l < M.n_layers; l++) {
        const int K = M.t_rows[l], N = M.t_n[l], rows = M.rows[l];
        const int outp = M.outp[l];
        const float *src = Wg + M.w_off[l];
        float *dst = smem_w + M.t_off[l];
        for (int i = threadIdx.x; i < K * N + N; i += blockDim.x) {
            const int k = i / N, j = i - k * N;
            float v = 0.0f;
            if (j < outp) {
                if (k < rows) v = src[k * outp + j];
                else if (k == K) v = src[rows * outp + j];   // the bias row
            }
            dst[i] = v;
        }
    }
    __syncthreads();
}

// deep_tracking_policy.py:241-286: fp64 differences first, then the fp32 cast
__device__ __forceinline__ void make_features(const double (&s)[12], const double (&tp)[3], const double (&tv)[3],
                                              float (&x)[QT_FEATURE_DIM]) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        x[i] = (float)(tp[i] - s[i]);
        x[3 + i] = (float)(tv[i] - s[3 + i]);
        x[6 + i] = (float)s[6 + i];
        x[9 + i] = (float)s[9 + i];
        x[12 + i] = (float)tp[i];
        x[15 + i] = (float)tv[i];
    }
}

// ---- tensor-core evaluation (qt_policy_tc.cuh) ------------------------------------------------------
// Stage the packed weights for the tensor-core evaluation: the hidden layers as B operands (unit n, input k) in the
// canonical K-major layout, split into TF32 head and tail; the output layer and the biases as plain fp32.
__device__ __forceinline__ void stage_weights_tc(const MlpDesc &M, const float *__restrict__ Wg, float *smem) {
    const tc::TcLayout &t = M.tc;
    for (int l = 0; l < t.hidden; l++) {
        const int K = t.k[l], N = t.n[l], rows = M.rows[l], outp = M.outp[l];
        const float *src = Wg + M.w_off[l];
        float *hi = smem + t.b_hi[l], *lo = smem + t.b_lo[l];
        for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
            const int k = i / N, n = i - k * N;
            const float w = (k < rows && n < outp) ? src[k * outp + n] : 0.0f;
            float h, tl;
            tc::split_tf32(w, h, tl);
            const int at = ((k >> 2) * N + n) * 4 + (k & 3);
            hi[at] = h;
            lo[at] = tl;
        }
        for (int n = threadIdx.x; n < tc::kN; n += blockDim.x)
            smem[t.bias + l * tc::kN + n] = n < outp ? src[rows * outp + n] : 0.0f;
    }
    {
        const int L = t.hidden, rows = M.rows[L], outp = M.outp[L];   // outp = 4
        const float *src = Wg + M.w_off[L];
        for (int i = threadIdx.x; i < tc::kN * 4; i += blockDim.x) {
            const int k = i >> 2, j = i & 3;
            smem[t.w3 + i] = k < rows ? src[k * outp + j] : 0.0f;
        }
        if (threadIdx.x < 4) smem[t.bias + tc::kMaxHidden * tc::kN + threadIdx.x] = src[rows * outp + threadIdx.x];
    }
}

// Second half of a hidden layer on the tensor cores: read this thread's n accumulators, add the bias, activate.
// NEXT_A: store the result as the next layer's A operand (heads to shared memory, tails to the set's TMEM columns);
// else multiply it into the four output accumulators.
// NFIX: the layer's width when it is known at compile time (the 64-64 default: fully unrolled), 0 = run time.
template <int KIND, bool NEXT_A, int NFIX>
__device__ __forceinline__ void tc_epilogue(const float *w3, float *a_hi, uint32_t taddr, const float *bias, int n_rt, int row,
                                            unsigned long long (&o)[2]) {
    constexpr int W = 16;   // accumulator columns per tcgen05.ld: quarters of a 64-wide row keep the live registers
                            // next to the fp64 env state low (32 columns at a time measured 6 % slower)
    const int n = NFIX ? NFIX : n_rt;
#pragma unroll
    for (int part = 0; part < (NFIX ? NFIX / W : tc::kN / W); part++) {
        if (!NFIX && part * W >= n) break;
        float d[W];
        [[maybe_unused]] float lo[W];
        tc::tmem_ld(taddr + part * W, d);
#pragma unroll
        for (int j = 0; j < W; j += 4) {
            const float4 b = *reinterpret_cast<const float4 *>(bias + part * W + j);
            const float h0 = activate_t<KIND>(d[j] + b.x), h1 = activate_t<KIND>(d[j + 1] + b.y);
            const float h2 = activate_t<KIND>(d[j + 2] + b.z), h3 = activate_t<KIND>(d[j + 3] + b.w);
            if (NEXT_A) {
                tc::store_chunk(a_hi, row, (part * W + j) >> 2, h0, h1, h2, h3, lo + j);
            } else {
                // output layer on the FP32 pipe: o[0] = (out0, out1), o[1] = (out2, out3) as FFMA2 pairs, ascending k
                const float h[4] = {h0, h1, h2, h3};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const ulonglong2 w = *reinterpret_cast<const ulonglong2 *>(w3 + (part * W + j + q) * 4);
                    ffma2(h[q], w.x, o[0]);
                    ffma2(h[q], w.y, o[1]);
                }
            }
        }
        if constexpr (NEXT_A) tc::tmem_st(taddr + 64 + part * W, lo);
    }
    if (NEXT_A) tc::tmem_st_wait();
}

// PolicyNetwork.forward for the 128 episodes of a set; every thread of the set must call it after having stored its
// feature row with tc_store_features and passed the set barrier that follows.
// FIX64x2: the reference's default shape, two hidden layers of 64 units — layer count and widths are compile-time
// constants (the run-time loops of the general form cost 7 % on that shape).
template <int KIND, bool FIX64x2>
__device__ __forceinline__ void mlp_forward_tc_kind(const MlpDesc &M, float *smem, uint32_t tmem_base, uint32_t &phase,
                                                    float (&out)[4]) {
    const tc::TcLayout &t = M.tc;
    constexpr int NF = FIX64x2 ? 64 : 0;
    const int hidden = FIX64x2 ? 2 : t.hidden;
    const int set = threadIdx.x / tc::kRows, row = threadIdx.x % tc::kRows;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + t.bar) + set;
    float *a_hi = smem + tc::kAHi + set * tc::kRows * tc::kK1;
    const uint32_t set_tmem = tmem_base + set * tc::kColsPerSet;
    const uint32_t taddr = set_tmem + ((uint32_t)(row & ~31) << 16);   // this warp's 32 TMEM lanes
    unsigned long long o[2];
    {
        const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(smem + t.bias + tc::kMaxHidden * tc::kN);
        o[0] = b.x;
        o[1] = b.y;
    }
#pragma unroll
    for (int l = 0; l < (FIX64x2 ? 2 : tc::kMaxHidden); l++) {
        if (!FIX64x2 && l >= hidden) break;
        if (row == 0) {
            tc::fence_after_sync();
            tc::issue_layer(smem, set, t.b_hi[l], t.b_lo[l], FIX64x2 ? 64 : t.n[l], FIX64x2 ? (l == 0 ? tc::kK0 / 8 : 8) : t.k[l] / 8,
                            set_tmem, bar);
        }
        tc::mbar_wait(bar, phase);
        phase ^= 1u;
        tc::fence_after_sync();
        if (l + 1 < hidden) {
            tc_epilogue<KIND, true, NF>(smem + t.w3, a_hi, taddr, smem + t.bias + l * tc::kN, t.n[l], row, o);
            tc::fence_before_sync();     // this thread's tcgen05.ld / st are done before the next MMAs touch those columns
            tc::fence_async_smem();      // ... and its new A heads are visible to the tensor core
            tc::set_sync(1 + set);
        } else {
            tc_epilogue<KIND, false, NF>(smem + t.w3, a_hi, taddr, smem + t.bias + l * tc::kN, t.n[l], row, o);
            tc::fence_before_sync();
        }
    }
    float of[4];
    unpack2(o[0], of[0], of[1]);
    unpack2(o[1], of[2], of[3]);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float sg = 1.0f / (1.0f + expf(-of[i]));  // torch.sigmoid
        out[i] = fmaf(sg, M.out_scale[i], M.out_lo[i]);
    }
}

template <bool FIX64x2>
__device__ __forceinline__ void mlp_forward_tc(const MlpDesc &M, float *smem, uint32_t tmem_base, uint32_t &phase,
                                               float (&out)[4]) {
    switch (M.activation) {
    case QT_ACT_TANH: mlp_forward_tc_kind<QT_ACT_TANH, FIX64x2>(M, smem, tmem_base, phase, out); break;
    case QT_ACT_ELU: mlp_forward_tc_kind<QT_ACT_ELU, FIX64x2>(M, smem, tmem_base, phase, out); break;
    case QT_ACT_LEAKY_RELU: mlp_forward_tc_kind<QT_ACT_LEAKY_RELU, FIX64x2>(M, smem, tmem_base, phase, out); break;
    default: mlp_forward_tc_kind<QT_ACT_RELU, FIX64x2>(M, smem, tmem_base, phase, out); break;
    }
}

// the 18 features of this thread's episode as chunks 0..5 of its A row (features 18..23 are zero padding)
__device__ __forceinline__ void tc_store_features(float *smem, uint32_t tmem_base, const float (&f)[QT_FEATURE_DIM]) {
    const int set = threadIdx.x / tc::kRows, row = threadIdx.x % tc::kRows;
    float *a_hi = smem + tc::kAHi + set * tc::kRows * tc::kK1;
    float lo[32];
#pragma unroll
    for (int i = 20; i < 32; i++) lo[i] = 0.0f;
    tc::store_chunk(a_hi, row, 0, f[0], f[1], f[2], f[3], lo + 0);
    tc::store_chunk(a_hi, row, 1, f[4], f[5], f[6], f[7], lo + 4);
    tc::store_chunk(a_hi, row, 2, f[8], f[9], f[10], f[11], lo + 8);
    tc::store_chunk(a_hi, row, 3, f[12], f[13], f[14], f[15], lo + 12);
    tc::store_chunk(a_hi, row, 4, f[16], f[17], 0.0f, 0.0f, lo + 16);
    *reinterpret_cast<float4 *>(a_hi + (5 * tc::kRows + row) * 4) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    tc::tmem_st(tmem_base + set * tc::kColsPerSet + 64 + ((uint32_t)(row & ~31) << 16), lo);
    tc::tmem_st_wait();
    tc::fence_before_sync();
    tc::fence_async_smem();
}

struct PolicyArgs {
    int64_t E, k_stride;
    int max_steps, w_in_smem, teacher;  // teacher: 0 none, else 1 + CtrlKind
    const double *state0, *tparams, *K;
    const float *W;
    double *metrics, *final_state, *traj;
    float *feat, *actf;
    int traj_stride, traj_slots, data_stride, data_slots;
};

// Stage the packed weights at the start of shared memory (when they fit).
__device__ __forceinline__ void stage_weights(const MlpDesc &M, const float *Wg, int in_smem, float *smem_w) {
    if (!in_smem) return;
    for (int i = threadIdx.x * 4; i < M.total_floats; i += blockDim.x * 4)
        *reinterpret_cast<float4 *>(smem_w + i) = *reinterpret_cast<const float4 *>(Wg + i);
    __syncthreads();
}

// MINB = resident blocks per SM the register allocator must allow.  Measured (2^18 episodes), per-thread
// evaluation: with the weights in shared memory three blocks of 168 registers beat two of 255 (64-64
// network 2.40e9 against 2.19e9 env-steps/s); when the weights are too large to stage and stream from
// L1/L2 instead, the deeper load prefetch that 255 registers allow wins (128-128: 4.0e8 against 2.4e8).
// TILED: the warp-cooperative evaluation (64 accumulators per lane, two blocks per SM).
// EVAL: 0 one episode per thread, 1 warp-cooperative tiles (both on the FP32 pipe, bit-identical), 2 / 3 the hidden
// layers on the tensor cores (qt_policy_tc.cuh; sets of 128 episodes are MMA operands, one 512-thread block per SM):
// 2 = the reference's default shape, two hidden layers of 64 units, with layer count and widths as compile-time
// constants; 3 = one to three hidden layers of any width up to 64 (run-time loops: 7 % slower on the 64-64 shape, and
// keeping both forms in ONE kernel cost the fixed form its register allocation, hence two instantiations).
template <int MOTION, int MINB, int EVAL>
__global__ void __launch_bounds__(EVAL >= 2 ? tc::kThreads : kPolicyThreads, MINB)
    rollout_policy_kernel(const __grid_constant__ qt_params P, const __grid_constant__ DynConsts C,
                          const __grid_constant__ MlpDesc M, const __grid_constant__ PolicyArgs A) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr bool TILED = EVAL == 1, TC = EVAL >= 2;
    constexpr int T = TC ? tc::kThreads : kPolicyThreads;
    const int64_t E = A.E;
    const int64_t e_raw = (int64_t)blockIdx.x * T + threadIdx.x;
    float *tile0 = nullptr, *tile1 = nullptr;
    uint32_t tmem_base = 0u, mma_phase = 0u;
    if constexpr (TC) {
        stage_weights_tc(M, A.W, smem_f);
        uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem_f + M.tc.tmem_slot);
        if (threadIdx.x < 32) tc::tmem_alloc(tmem_slot);
        if (threadIdx.x % tc::kRows == 0) tc::mbar_init(reinterpret_cast<uint64_t *>(smem_f + M.tc.bar) + threadIdx.x / tc::kRows, 1);
        tc::fence_async_smem();      // the staged weights are read by the tensor core
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        tmem_base = *tmem_slot;
    } else if constexpr (TILED) {
        stage_weights_tiled(M, A.W, smem_f);
        float *wb = smem_f + M.t_total + (threadIdx.x >> 5) * (M.t_buf[0] + M.t_buf[1]) * 32;
        tile0 = wb;
        tile1 = wb + M.t_buf[0] * 32;
    } else {
        const int w_floats = A.w_in_smem ? ((M.total_floats + 3) & ~3) : 0;
        stage_weights(M, A.W, A.w_in_smem, smem_f);
        tile0 = smem_f + w_floats + threadIdx.x;
        tile1 = tile0 + M.tile_w[0] * T;
        if (e_raw >= E) return;
    }
    // TILED / TC: the lanes past the last episode stay, because the whole warp (block) evaluates the network
    // together; they shadow the last episode and never write
    const bool valid = e_raw < E;
    const int64_t e = valid ? e_raw : E - 1;

    double s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = A.state0[(int64_t)i * E + e];
    double par[QT_TARGET_PARAMS];
    TargetGen<MOTION, 1> tg;
    tg.load(par, A.tparams, E, e);

    // supervisor state: reset before every episode (train.py:665-667)
    const int teacher_kind = A.teacher - 1;
    const int KC = teacher_kind == CK_LQI ? 9 : (teacher_kind == CK_PID ? 3 : 6);
    const double *Kp = A.teacher ? A.K + (e / A.k_stride) * 4 * KC : nullptr;
    double integ[3] = {0.0, 0.0, 0.0};
    double last_time = __longlong_as_double(0x7ff8000000000000LL);
    int saturations = 0;

    double t = 0.0;
    int step = 0, env_on = 0, met_on = 0, violations = 0, term = QT_TERM_NONE;
    double err_post = 0.0, sum_err = 0.0, sum_err2 = 0.0, sum_eff = 0.0, max_err = 0.0, ov_cur = 0.0, ov_max = 0.0;
    bool prev_on = false, in_phase = false;
    int streak = 0, ov_count = 0;
    int next_traj = 0, traj_slot = 0, next_data = 0, data_slot = 0;

    double tp[3], tv[3], ta[3];
    tg.template state<true>(P, t, tp, tv, ta);
    double err_pre = norm3d(tp[0] - s[0], tp[1] - s[1], tp[2] - s[2]);

    bool alive = valid && A.max_steps > 0;
    while (true) {
        float feat[QT_FEATURE_DIM];
        if constexpr (TC) {
            // the 128 episodes of a set leave together; the barrier that takes the vote also publishes the feature rows
            make_features(s, tp, tv, feat);
            tc_store_features(smem_f, tmem_base, feat);
            if (!tc::set_sync_or(1 + threadIdx.x / tc::kRows, alive)) break;
        } else if constexpr (TILED) {
            if (!__any_sync(0xffffffffu, alive)) break;
        } else {
            if (!alive) break;
        }
        // ---- the policy acts on the pre-step observation
        if constexpr (!TC) make_features(s, tp, tv, feat);
        const bool log_data = alive && A.feat && step == next_data && data_slot < A.data_slots;
        if (log_data) {
            float *f = A.feat + ((int64_t)data_slot * E + e) * QT_FEATURE_DIM;
#pragma unroll
            for (int i = 0; i < QT_FEATURE_DIM; i += 2) *reinterpret_cast<float2 *>(f + i) = make_float2(feat[i], feat[i + 1]);
        }
        float out[4];
        if constexpr (TC)
            mlp_forward_tc<EVAL == 2>(M, smem_f, tmem_base, mma_phase, out);
        else if constexpr (TILED)
            mlp_forward_tiled(M, smem_f, tile0, tile1, feat, out);
        else if (A.w_in_smem)   // two inlined copies: one reads the weights with LDS, the other with LDG
            mlp_forward<T>(M, smem_f, tile0, tile1, feat, out);
        else
            mlp_forward<T>(M, A.W, tile0, tile1, feat, out);
        if (!alive) continue;   // TILED / TC only: this lane's episode is over, it only helps the warp / block
        double act[4] = {(double)out[0], (double)out[1], (double)out[2], (double)out[3]};

        // ---- supervisor label on the same observation (train.py:682-696)
        double lab[4] = {act[0], act[1], act[2], act[3]};
        if (A.teacher) {
            if (teacher_kind == CK_PID)
                pid_action<true>(P, Kp, s, tp, tv, ta, t, integ, last_time, lab);
            else if (teacher_kind == CK_LQI)
                ctrl_action<true, true>(P, Kp, s, tp, tv, ta, err_pre, integ, saturations, lab);
            else
                ctrl_action<false, true>(P, Kp, s, tp, tv, ta, err_pre, integ, saturations, lab);
        }

        // ---- metrics sample k = step (utils/metrics.py:264-338 streamed, as in qt_rollout.cu)
        {
            const bool on = err_pre <= P.crit_target_radius;
            met_on += on ? 1 : 0;
            sum_err += err_pre;
            sum_err2 += err_pre * err_pre;
            if (step == 0 || err_pre > max_err) max_err = err_pre;
            sum_eff += sqrt(act[0] * act[0] + act[1] * act[1] + act[2] * act[2] + act[3] * act[3]);
            if (step > 0) {
                if (prev_on && !on) {
                    in_phase = true;
                    streak = 1;
                    ov_cur = err_pre - P.crit_target_radius;
                } else if (in_phase && !on) {
                    streak += 1;
                    ov_cur = fmax(ov_cur, err_pre - P.crit_target_radius);
                } else if (in_phase && on) {
                    if (streak >= P.overshoot_window) {
                        ov_count += 1;
                        ov_max = fmax(ov_max, ov_cur);
                    }
                    in_phase = false;
                    streak = 0;
                    ov_cur = 0.0;
                }
            }
            prev_on = on;
        }

        bool log_traj = false;
        if (A.traj && step == next_traj && traj_slot < A.traj_slots) {
            log_traj = true;
            double *row = A.traj + (int64_t)traj_slot * QT_TRAJ_FIELDS * E + e;
#pragma unroll
            for (int i = 0; i < 12; i++) row[(int64_t)(QT_TRAJ_STATE + i) * E] = s[i];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                row[(int64_t)(QT_TRAJ_TARGET_P + i) * E] = tp[i];
                row[(int64_t)(QT_TRAJ_TARGET_V + i) * E] = tv[i];
            }
#pragma unroll
            for (int i = 0; i < 4; i++) row[(int64_t)(QT_TRAJ_ACTION + i) * E] = act[i];
        }
        if (log_data) {
            float2 *y = reinterpret_cast<float2 *>(A.actf + ((int64_t)data_slot * E + e) * QT_ACTION_DIM);
            y[0] = make_float2((float)lab[0], (float)lab[1]);
            y[1] = make_float2((float)lab[2], (float)lab[3]);
            next_data += A.data_stride;
            data_slot += 1;
        }

        // ---- env.step (quadcopter_env.py:171-207)
        double actv[4];
        violations += validate_action<true>(P, act, actv);
        if (P.integrator == QT_INTEGRATOR_EULER)
            integrate<true>(C, s, actv);
        else
            integrate<false>(C, s, actv);
        constrain_state(P, C, s);
        t += C.dt;
        tg.template state<true>(P, t, tp, tv, ta);
        err_post = norm3d(s[0] - tp[0], s[1] - tp[1], s[2] - tp[2]);
        term = termination_code(P, C, s, t);
        env_on += (err_post <= P.env_target_radius) ? 1 : 0;
        if (log_traj) {
            double *row = A.traj + (int64_t)traj_slot * QT_TRAJ_FIELDS * E + e;
            row[(int64_t)QT_TRAJ_TIME * E] = t;
            row[(int64_t)QT_TRAJ_ERR * E] = err_post;
            next_traj += A.traj_stride;
            traj_slot += 1;
        }
        step += 1;
        err_pre = err_post;
        alive = term == QT_TERM_NONE && step < A.max_steps;
    }
    if constexpr (TC) {
        tc::fence_before_sync();
        __syncthreads();
        if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base);
    }
    if (!valid) return;
    if (in_phase && streak >= P.overshoot_window) {
        ov_count += 1;
        ov_max = fmax(ov_max, ov_cur);
    }
    const double n = (double)step;
    const double env_ratio = step > 0 ? (double)env_on / n : 0.0;
    const double met_ratio = step > 0 ? (double)met_on / n : 0.0;
    double *m = A.metrics + e;
    m[(int64_t)QT_M_STEPS * E] = n;
    m[(int64_t)QT_M_FINAL_TIME * E] = t;
    m[(int64_t)QT_M_ENV_ON_TARGET * E] = (double)env_on;
    m[(int64_t)QT_M_MET_ON_TARGET * E] = (double)met_on;
    m[(int64_t)QT_M_SUM_ERR * E] = sum_err;
    m[(int64_t)QT_M_MAX_ERR * E] = max_err;
    m[(int64_t)QT_M_SUM_ERR2 * E] = sum_err2;
    m[(int64_t)QT_M_SUM_EFFORT * E] = sum_eff;
    m[(int64_t)QT_M_OVERSHOOT_COUNT * E] = (double)ov_count;
    m[(int64_t)QT_M_MAX_OVERSHOOT * E] = ov_max;
    m[(int64_t)QT_M_ENV_SUCCESS * E] =
        (t >= P.env_min_episode_duration && env_ratio >= P.env_min_on_target_ratio) ? 1.0 : 0.0;
    m[(int64_t)QT_M_MET_SUCCESS * E] =
        (step > 0 && t >= P.crit_min_episode_duration && met_ratio >= P.crit_min_on_target_ratio) ? 1.0 : 0.0;
    m[(int64_t)QT_M_TERMINATION * E] = (double)term;
    m[(int64_t)QT_M_VIOLATIONS * E] = (double)violations;
    m[(int64_t)QT_M_SATURATIONS * E] = (double)saturations;
    m[(int64_t)QT_M_FINAL_ERR * E] = err_post;
    if (A.final_state) {
#pragma unroll
        for (int i = 0; i < 12; i++) A.final_state[(int64_t)i * E + e] = s[i];
    }
}

struct PolicyActionArgs {
    int64_t E;
    int w_in_smem;
    const double *obs;
    const float *W;
    double *action;
};

__global__ void __launch_bounds__(kPolicyThreads)
    policy_action_kernel(const __grid_constant__ MlpDesc M, const __grid_constant__ PolicyActionArgs A) {
    extern __shared__ __align__(16) float smem_f[];
    constexpr int T = kPolicyThreads;
    const int64_t E = A.E;
    const int64_t e = (int64_t)blockIdx.x * T + threadIdx.x;
    const int w_floats = A.w_in_smem ? ((M.total_floats + 3) & ~3) : 0;
    stage_weights(M, A.W, A.w_in_smem, smem_f);
    float *tile0 = smem_f + w_floats + threadIdx.x;
    float *tile1 = tile0 + M.tile_w[0] * T;
    if (e >= E) return;
    double s[12], tp[3], tv[3];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = A.obs[(int64_t)i * E + e];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        tp[i] = A.obs[(int64_t)(12 + i) * E + e];
        tv[i] = A.obs[(int64_t)(15 + i) * E + e];
    }
    float feat[QT_FEATURE_DIM];
    make_features(s, tp, tv, feat);
    float out[4];
    if (A.w_in_smem)
        mlp_forward<T>(M, smem_f, tile0, tile1, feat, out);
    else
        mlp_forward<T>(M, A.W, tile0, tile1, feat, out);
#pragma unroll
    for (int i = 0; i < 4; i++) A.action[(int64_t)i * E + e] = (double)out[i];
}

// shared memory plan, per-thread evaluation: [weights if they fit][activation tile 0][activation tile 1]
static size_t smem_plan(const MlpDesc &M, int &w_in_smem) {
    const size_t act = (size_t)(M.tile_w[0] + M.tile_w[1]) * kPolicyThreads * sizeof(float);
    const size_t w = (((size_t)M.total_floats + 3) & ~(size_t)3) * sizeof(float);
    const size_t budget = 100 * 1024;  // at least two blocks per SM
    w_in_smem = (act + w <= budget) ? 1 : 0;
    return act + (w_in_smem ? w : 0);
}

// tiled evaluation: [staged weights][per warp: buffer 0, buffer 1]
static size_t smem_plan_tiled(const MlpDesc &M) {
    return ((size_t)M.t_total + (size_t)(kPolicyThreads / 32) * (M.t_buf[0] + M.t_buf[1]) * 32) * sizeof(float);
}

template <int MOTION, int MINB, int EVAL>
static cudaError_t launch_policy_as(const qt_params &P, const DynConsts &C, const MlpDesc &M, const PolicyArgs &A,
                                    size_t smem, cudaStream_t st) {
    constexpr int threads = EVAL >= 2 ? tc::kThreads : kPolicyThreads;
    const int64_t blocks = (A.E + threads - 1) / threads;
    cudaError_t e = cudaFuncSetAttribute(rollout_policy_kernel<MOTION, MINB, EVAL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rollout_policy_kernel<MOTION, MINB, EVAL><<<(unsigned)blocks, threads, smem, st>>>(P, C, M, A);
    return cudaGetLastError();
}

template <int MOTION>
static cudaError_t launch_policy(const qt_params &P, const MlpDesc &M, PolicyArgs &A, cudaStream_t st, bool &too_wide) {
    DynConsts C;
    C.init(P);
    too_wide = false;
    if (M.tc_ok && (P.debug_flags & QT_FLAG_POLICY_TENSOR_CORES) && !(P.debug_flags & QT_DEBUG_POLICY_PER_THREAD))
    {
        const bool fix = M.tc.hidden == 2 && M.tc.n[0] == 64 && M.tc.n[1] == 64;
        if (fix) return launch_policy_as<MOTION, 1, 2>(P, C, M, A, (size_t)M.tc.floats * sizeof(float), st);
        return launch_policy_as<MOTION, 1, 3>(P, C, M, A, (size_t)M.tc.floats * sizeof(float), st);
    }
    if (M.tiled && !(P.debug_flags & QT_DEBUG_POLICY_PER_THREAD) && smem_plan_tiled(M) <= 227 * 1024) {
        // three resident blocks (168 registers, the env state spills to L1 across the network evaluation)
        // when the shared memory allows, else two or one
        if (QT_POLICY_TILED_3 && smem_plan_tiled(M) <= 74 * 1024)
            return launch_policy_as<MOTION, 3, 1>(P, C, M, A, smem_plan_tiled(M), st);
        return launch_policy_as<MOTION, 2, 1>(P, C, M, A, smem_plan_tiled(M), st);
    }
    const size_t smem = smem_plan(M, A.w_in_smem);
    if (smem > 200 * 1024) {
        too_wide = true;
        return cudaSuccess;
    }
    if (A.w_in_smem && smem <= 72 * 1024) return launch_policy_as<MOTION, 3, 0>(P, C, M, A, smem, st);
    return launch_policy_as<MOTION, 2, 0>(P, C, M, A, smem, st);
}

}  // namespace qt

extern "C" int64_t qt_mlp_packed_floats(const qt_mlp *policy) {
    qt::MlpDesc d;
    if (!policy || !qt::describe(*policy, d)) return -1;
    return d.total_floats;
}

extern "C" int qt_policy_action_batched(qt_context *ctx, const qt_mlp *policy, int64_t E, const double *obs_dev,
                                        double *action_out_dev) {
    using namespace qt;
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, policy && obs_dev && action_out_dev && policy->weights_dev, "null pointer");
    QT_REQUIRE(ctx, E >= 0, "E < 0");
    MlpDesc M;
    QT_REQUIRE(ctx, describe(*policy, M), "unsupported network shape (18 inputs, 4 outputs, <= 6 layers, widths <= 256)");
    QT_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(policy->weights_dev) & 15) == 0, "weights must be 16-byte aligned");
    if (E == 0) return QT_OK;
    PolicyActionArgs A{E, 0, obs_dev, policy->weights_dev, action_out_dev};
    const size_t smem = smem_plan(M, A.w_in_smem);
    QT_REQUIRE(ctx, smem <= 200 * 1024, "network too wide for the shared-memory activation tiles");
    DeviceGuard guard(ctx->device);
    QT_CUDA_TRY(ctx, cudaFuncSetAttribute(policy_action_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = (E + kPolicyThreads - 1) / kPolicyThreads;
    policy_action_kernel<<<(unsigned)blocks, kPolicyThreads, smem, ctx->stream>>>(M, A);
    return check_launch(ctx, "policy_action_kernel");
}

extern "C" int qt_rollout_policy(qt_context *ctx, const qt_params *params, int64_t E, int64_t max_steps,
                                 const double *state0_dev, const double *tparams_dev, const qt_mlp *policy,
                                 const double *teacher_K_dev, int32_t k_cols, int64_t k_stride,
                                 double *metrics_out_dev, double *final_state_out_dev, double *traj_out_dev,
                                 int64_t traj_stride, int64_t traj_slots, float *feat_out_dev, float *act_out_dev,
                                 int64_t data_stride, int64_t data_slots) {
    using namespace qt;
    if (!ctx) return QT_ERR_ARG;
    QT_REQUIRE(ctx, params && state0_dev && tparams_dev && policy && policy->weights_dev && metrics_out_dev, "null pointer");
    QT_REQUIRE(ctx, E >= 0 && max_steps >= 0 && max_steps < (1ll << 31), "bad E / max_steps");
    QT_REQUIRE(ctx, params->motion_type >= 0 && params->motion_type <= QT_MOTION_FIGURE8, "motion_type");
    QT_REQUIRE(ctx, (feat_out_dev == nullptr) == (act_out_dev == nullptr), "feat/act must come together");
    MlpDesc M;
    QT_REQUIRE(ctx, describe(*policy, M), "unsupported network shape (18 inputs, 4 outputs, <= 6 layers, widths <= 256)");
    QT_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(policy->weights_dev) & 15) == 0, "weights must be 16-byte aligned");
    if (teacher_K_dev) {
        QT_REQUIRE(ctx, k_cols == expected_k_cols(*params), "k_cols must be 6 (LQR), 9 (LQI) or 3 (PID)");
        QT_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(teacher_K_dev) & 15) == 0, "K must be 16-byte aligned");
    }
    if (E == 0) return QT_OK;
    if (k_stride < 1) k_stride = 1;
    if (traj_stride < 1) traj_stride = 1;
    if (data_stride < 1) data_stride = 1;
    QT_REQUIRE(ctx, traj_stride < (1ll << 31) && data_stride < (1ll << 31) && traj_slots < (1ll << 31) &&
                        data_slots < (1ll << 31), "log strides/slots");
    PolicyArgs A;
    A.E = E;
    A.k_stride = k_stride;
    A.max_steps = (int)max_steps;
    A.teacher = teacher_K_dev ? 1 + ctrl_kind_of(*params) : 0;
    A.state0 = state0_dev;
    A.tparams = tparams_dev;
    A.K = teacher_K_dev;
    A.W = policy->weights_dev;
    A.metrics = metrics_out_dev;
    A.final_state = final_state_out_dev;
    A.traj = traj_out_dev;
    A.feat = feat_out_dev;
    A.actf = act_out_dev;
    A.traj_stride = (int)traj_stride;
    A.traj_slots = (int)traj_slots;
    A.data_stride = (int)data_stride;
    A.data_slots = (int)data_slots;
    A.w_in_smem = 0;
    DeviceGuard guard(ctx->device);
    cudaError_t err;
    bool too_wide = false;
    switch (params->motion_type) {
    case QT_MOTION_LINEAR: err = launch_policy<QT_MOTION_LINEAR>(*params, M, A, ctx->stream, too_wide); break;
    case QT_MOTION_CIRCULAR: err = launch_policy<QT_MOTION_CIRCULAR>(*params, M, A, ctx->stream, too_wide); break;
    case QT_MOTION_SINUSOIDAL: err = launch_policy<QT_MOTION_SINUSOIDAL>(*params, M, A, ctx->stream, too_wide); break;
    case QT_MOTION_FIGURE8: err = launch_policy<QT_MOTION_FIGURE8>(*params, M, A, ctx->stream, too_wide); break;
    default: err = launch_policy<QT_MOTION_STATIONARY>(*params, M, A, ctx->stream, too_wide); break;
    }
    QT_REQUIRE(ctx, !too_wide, "network too wide for the shared-memory activation tiles");
    if (err != cudaSuccess) return fail(ctx, QT_ERR_CUDA, "rollout_policy_kernel launch", cudaGetErrorString(err));
    ctx->launches += 1;
    return QT_OK;
}
