// qt_policy_tc.cuh — the two hidden layers of the deep tracking policy on the 5th-generation tensor cores
// (controllers/deep_tracking_policy.py:114-147, PolicyNetwork.forward).
//
// A "set" of 128 threads = 128 episodes is one M = 128 operand: episode i is row i of A and lane i of the
// accumulator in tensor memory.  A block is kSets independent sets (own mbarrier, own named barrier, own TMEM columns)
// that share one staged copy of the weights — the weights (44 KB) are what kept a one-set block at two blocks = 8 warps
// per SM; four sets in one block are 16 warps per SM.  Per env step and layer:
//   every thread splits its activation row into a TF32 head and tail (x = hi + lo, hi = cvt.rna.tf32(x),
//   lo = x - hi exactly); the heads go to shared memory in the canonical K-major core-matrix layout (no swizzle:
//   8 rows x 16 bytes per core matrix; the 16-byte chunk c of row m lives at (c * 128 + m) * 16, so a warp's 16-byte
//   stores are consecutive and conflict free), the tails to the set's TMEM columns with `tcgen05.st` (the thread's
//   own lane: no shared memory, which is what makes room for four sets);
//   one thread of the set issues, per k-step of 8, the three `tcgen05.mma.cta_group::1.kind::tf32` of the 3xTF32
//   scheme — A_lo[TMEM] B_hi + A_hi[smem] B_lo + A_hi[smem] B_hi; the weights were split once when the block staged
//   them; fp32 accumulation in TMEM — and a `tcgen05.commit` onto the set's mbarrier;
//   every thread waits on the mbarrier, reads its own 64 accumulators with `tcgen05.ld.32x32b` (warp w owns TMEM
//   lanes 32 (w % 4) .. + 31), adds the bias and applies the activation.
// The 64 -> 4 output layer stays on the FP32 pipe, from registers (256 FMAs per episode: a tensor-core round trip
// would cost more than it saves, and this keeps torch's order of summation for that layer).
// Accuracy: dropping lo * lo and the truncation of lo to 11 bits leave a relative error of about 2^-21 per product —
// the size of fp32 summation-order noise in a 64-term dot product — so this evaluation is NOT bit-identical to the
// FP32-pipe evaluations (qt_policy.cu), which remain the default and the parity path; the tests bound the difference.
// Shapes: one to three hidden layers of at most 64 units each (the reference's default 64-64, its "fast" 32-32), every
// width zero padded to a multiple of 16 (a padded unit has zero weights and bias, so it outputs activation(0) = 0).
// Shared memory per block: kSets x 32 KB (A heads, K = 64) + the split weights (64-64: 44 KB) + 2 KB = 174 KB with
// four sets: one block of 512 threads per SM.  TMEM: 128 columns per set (64 accumulators + 64 tails).
#pragma once
#include <stdint.h>

namespace qt {
namespace tc {

#ifndef QT_TC_SETS
#define QT_TC_SETS 4
#endif
constexpr int kSets = QT_TC_SETS;          // independent 128-episode sets per block
constexpr int kRows = 128;                 // episodes per set = M of the MMA
constexpr int kThreads = kSets * kRows;
constexpr int kN = 64;                     // largest hidden width (a layer's width is padded to a multiple of 16)
constexpr int kMaxHidden = 3;              // hidden layers on the tensor cores
constexpr int kK0 = 24;                    // 18 features padded to a multiple of the MMA's K = 8
constexpr int kK1 = 64;                    // largest K of a later layer = largest width
constexpr int kColsPerSet = 128;           // TMEM columns of a set: accumulator [0, 64) | A tails [64, 128)
constexpr int kTmemCols = kSets * kColsPerSet <= 128 ? 128 : (kSets * kColsPerSet <= 256 ? 256 : 512);
static_assert(kSets >= 1 && kSets <= 4, "512 TMEM columns hold four sets");
// Shared memory of a block (float offsets).  The A heads come first; what follows depends on the network's shape
// and is laid out on the host (TcLayout, part of the kernel's MlpDesc argument).
constexpr int kAHi = 0;                                  // [set][16 chunks][128 rows][4]
constexpr int kWeights = kAHi + kSets * kRows * kK1;     // B operands of the hidden layers, hi then lo per layer

struct TcLayout {
    int hidden;                  // hidden layers evaluated on the tensor cores (1..kMaxHidden)
    int n[kMaxHidden];           // layer widths, padded to multiples of 16
    int k[kMaxHidden];           // layer depths: 24 for the first, the previous width for the others
    int b_hi[kMaxHidden];        // float offsets of the split weights [K / 4 chunks][N units][4]
    int b_lo[kMaxHidden];
    int w3;                      // output layer [64][4] fp32
    int bias;                    // hidden biases [kMaxHidden][64], then the 4 output biases
    int bar;                     // kSets mbarriers (8 bytes each, 8-byte aligned)
    int tmem_slot;               // TMEM base address written by tcgen05.alloc
    int floats;                  // total
};

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2),
// both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t idesc_for(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
}


__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), no swizzle, K-major: start address >> 4 at bit 0,
// leading byte offset (between the two 16-byte K chunks of one MMA) >> 4 at bit 16, stride byte offset (between
// groups of 8 rows) >> 4 at bit 32, descriptor version 1 at bit 46
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// the same with the A operand in tensor memory (row i = lane i, one 32-bit column per K element)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// named barrier of one set (id 1 + set, 128 threads); the second form also ORs a predicate over the set
__device__ __forceinline__ void set_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ bool set_sync_or(int id, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.u32 q, %2, 0;\n\t"
        "bar.red.or.pred p, %1, 128, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(r)
        : "r"(id), "r"((uint32_t)pred)
        : "memory");
    return r != 0u;
}

__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem) {   // one whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_dealloc(uint32_t base) {      // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(kTmemCols) : "memory");
}

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// 32 consecutive columns of this thread's TMEM lane from registers (complete after tmem_st_wait)
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[32]) {
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = __float_as_uint(v[i]);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 16-column forms (the epilogues work on quarters of a row: fewer live registers next to the fp64 env state)
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[16]) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = __float_as_uint(v[i]);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// x = hi + lo with hi the nearest TF32 (ties away) and lo = x - hi exact in fp32 (the MMA reads lo's top 19 bits)
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    lo = x - hi;
}

// four consecutive K values of this thread's row (chunk c): heads to the set's A tile in shared memory, tails to lo[]
__device__ __forceinline__ void store_chunk(float *a_hi, int row, int c, float x0, float x1, float x2, float x3, float *lo) {
    float4 hi;
    split_tf32(x0, hi.x, lo[0]);
    split_tf32(x1, hi.y, lo[1]);
    split_tf32(x2, hi.z, lo[2]);
    split_tf32(x3, hi.w, lo[3]);
    *reinterpret_cast<float4 *>(a_hi + (c * kRows + row) * 4) = hi;
}

// one layer's MMAs of one set: D[128][N] = A[128][K] B[N][K]', three TF32 products per k-step of 8; one thread
__device__ __forceinline__ void issue_layer(const float *smem, int set, int b_hi, int b_lo, int n, int ksteps,
                                            uint32_t set_tmem, uint64_t *bar) {
    const uint32_t a_hi = smem_u32(smem + kAHi + set * kRows * kK1);
    const uint32_t bh = smem_u32(smem + b_hi), bl = smem_u32(smem + b_lo);
    constexpr uint32_t kALbo = kRows * 16, kSbo = 128;
    const uint32_t b_lbo = (uint32_t)n * 16, idesc = idesc_for(n);
    const uint32_t d_tmem = set_tmem, a_lo = set_tmem + 64;
    for (int ks = 0; ks < ksteps; ks++) {
        const uint64_t dah = smem_desc(a_hi + ks * 2 * kALbo, kALbo, kSbo);
        const uint64_t dbh = smem_desc(bh + ks * 2 * b_lbo, b_lbo, kSbo), dbl = smem_desc(bl + ks * 2 * b_lbo, b_lbo, kSbo);
        mma_tf32_ts(d_tmem, a_lo + ks * 8, dbh, idesc, ks > 0 ? 1u : 0u);
        mma_tf32(d_tmem, dah, dbl, idesc, 1u);
        mma_tf32(d_tmem, dah, dbh, idesc, 1u);
    }
    mma_commit(bar);
}

}  // namespace tc
}  // namespace qt
