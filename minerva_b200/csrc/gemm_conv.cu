// gemm_conv.cu -- a18 MatMult and a14-a16 Convolution forward / backward-data / backward-filter as
// ONE warp-specialised tcgen05 kernel: TF32 operands, fp32 accumulation in TMEM.
//
// Replaces cublasSgemm (L1, minerva/op/impl/cuda/cuda_perform.cu:66-70) and the cuDNN-v2 convolution
// calls L7-L9 (cuda_perform.cu:227-318: per-call descriptor churn, algorithm search, cudaMalloc /
// cudaFree of the workspace, a separate bias pass and a stream sync).  Here every op is an
// enqueue-only launch; bias is fused into the epilogue.
//
// Formulation.  Every op is D[M x N] = sum_k A[m,k] * B[n,k] with both operands staged K-major in
// shared memory (UMMA canonical layout, 128-byte swizzle), D accumulated in TMEM (128 lanes x bn
// fp32 columns, double buffered) and written back so that the lane (= row m) index runs along the
// contiguous axis of the output:
//   MatMult        m = row of C,            n = column of C,  k = inner;   A col-major (m-contig.)
//   ConvForward    m = (img, oh, ow),       n = co,           k = (ci,r,s) A = im2col gather of x
//   ConvBackwardData   m = (img, h, w),     n = ci,           k = (co,r,s) A = gather of top_diff
//   ConvBackwardFilter m = (ci,r,s),        n = co,           k = (img,oh,ow)  A = gather of x
// (r,s) index the filter as stored; the 180-degree rotation of CUDNN_CONVOLUTION is the
// identity kh = fh-1-r, kw = fw-1-s applied inside the gathers.
//
// NCHW has no TMA-friendly im2col, so operands are gathered by producer warps with coalesced
// LDG (lanes run along the contiguous axis of the source), rounded to TF32 (cvt.rna), and stored
// with conflict-free 128-bit STS into the swizzled tile; fence.proxy.async + mbarrier hands the
// stage to the single MMA-issuing thread.  Warp roles (704 threads, 1 CTA/SM, persistent over tiles):
//   warps 0-3  epilogue: tcgen05.ld 32x32b -> +bias -> coalesced STG (lane = row m)
//   warp  4    TMEM alloc/dealloc; lane 0 issues tcgen05.mma.cta_group::1.kind::tf32 + tcgen05.commit
//   warp  5    TMA producer for B when B is a plain K-major matrix (cp.async.bulk.tensor, 128B swizzle)
//   warps 6-21 gather producers (512 threads = 4 warps per SM sub-partition: the gathers are
//              issue-bound, so thread-level parallelism matters): A tile 128x32 (+ B when it is not
//              TMA-able) per stage, 4-stage ring, 2-stage register look-ahead
// Split-K (needs the caller's workspace) keeps all 148 SMs busy when M*N has few tiles (FC layers,
// backward-filter); partials are folded in split order by a second kernel => deterministic.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string>
#include "common.cuh"

namespace mnv {

// ------------------------------------------------------------------------------------------------
// problem description shared by the tcgen05 kernel, the SIMT checker and the split-K reducer
// ------------------------------------------------------------------------------------------------
enum : int { A_COLMAJOR = 0, A_IM2COL_FWD = 1, A_IM2COL_BWD = 2, A_IM2COL_WGRAD = 3, A_TMA = 4,
              // Strided convolutions (first layers: 11x11/4, 7x7/2 on 3 channels): along the output pixels the source
              // addresses are `stride` elements apart, so lanes-along-pixels gathers touch 16 sectors per request.
              // These two modes put the lanes along the filter taps instead, which are contiguous along kw:
              A_IM2COL_FWD_K = 5,     // forward: a warp owns 32 tile rows, lane = k of the stage, one row per load
              A_IM2COL_WGRAD_M = 6 }; // backward-filter: a warp owns 32 tile rows (taps), lane = row, one pixel per load
// A_TMA sub-modes (GemmParams::a_mode): how the TMA thread fetches the 128 x 32 A tile of a k-stage
//   TMA_A_IM2COL_K : channels-last activation copy through an im2col tensor map; one box of 128 output pixels
//                    x 32 channels of one filter tap => K-major tile (forward conv, stride-1 backward-data)
//   TMA_A_TILED_MN : column-major matrix; 4 boxes of 32 m x 32 k => MN-major tile (MatMult)
//   TMA_A_IM2COL_MN: im2col map again, but the 32 pixels are the k axis and 4 boxes of 32 channels (each box its
//                    own (tap, channel chunk)) the m axis => MN-major tile (backward-filter)
//   TMA_A_TILED_K  : row-major (transposed) matrix: one 128-row x 32-k box per UMMA half => K-major tile (MatMult, A^T)
enum : int { TMA_A_IM2COL_K = 1, TMA_A_TILED_MN = 2, TMA_A_IM2COL_MN = 3, TMA_A_TILED_K = 4 };
enum : int { B_KMAJOR = 0, B_DY_WGRAD = 1,
              // backward-data without a workspace: B[n = ci][k = (co, r, s)] read straight out of the filter as stored,
              // filter[co][ci][r][s] (taps rotated by 180 degrees when GemmParams::b_flip) -- no re-laid-out copy
              B_FILTER_T = 2 };

struct GemmParams {
  const float* a;
  const float* b;
  const float* bias;   // per-n, may be null
  float* out;
  float* partial;      // split-K partials [split][n][m], null when splits == 1
  int M, N, K;
  int lda;             // A_COLMAJOR leading dimension
  int ldb, b_vec;      // B_KMAJOR row pitch (floats); 1 if rows are 16B-aligned and K % 4 == 0
  // convolution geometry (bottom H x W, top Ho x Wo)
  int Ci, Co, H, W, Ho, Wo, fh, fw, ph, pw, sv, sh;
  // epilogue addressing: out[(m / P) * img_stride + (m % P) + n * col_stride]
  int P;
  long long img_stride, col_stride;
  // tiling
  int bn, m_tiles, n_tiles, splits, stages_per_split, k_stages;
  // Tail split (splits == 1 only): a tile grid that ends in a thin last wave gives that wave's tiles to tail_splits CTAs
  // each (partials + tail reduce), so the machine stays full -- 338 tiles on 148 SMs cost 2 + 1/3 waves instead of 3.
  int tail_first;       // first tile id (raster order) that is K-split; == m_tiles * n_tiles when there is no tail
  int tail_splits, tail_stages;
  int pf_dist;          // L2 prefetch distance in k-stages for the TMA-fed operands (0 = none)
  int total_units;      // work items of the persistent loop: tail_first + (tiles - tail_first) * tail_splits, or tiles * splits
  int use_ktab;         // A_IM2COL_FWD: k -> (offset, kh, kw) table in shared memory
  int spi;              // backward-filter with TMA-fed top_diff: k-stages per image (K padded per image), else 0
  unsigned wait_hint;   // mbarrier.try_wait suspend-time hint (ns)
  int stages;           // smem ring depth: 4 (bn <= 256) or 3 (wide tile, 256 < bn <= 384)
  int stage_bytes;      // A tile + B tile bytes per ring slot
  int wide;             // 1: one 128 x bn tile as two UMMA halves of bn/2 columns sharing the A tile, single accumulator
  int tall;             // 1: 256 x bn tile as two UMMA halves of 128 rows sharing the B tile (all-TMA path only)
  int a_mode;           // A_TMA sub-mode (TMA_A_*), 0 otherwise
  int relu;             // 1: epilogue applies max(x, 0) after the bias (conv + ReLU units fused by owl.net)
  int b_mn;             // 1: B arrives as an MN-major tile (bn/32 boxes of 32 n x 32 k; MatMult with B^T), all-TMA path only
                        // 2: the same tile as ONE box of a (32 n, k, n / 32) view of the matrix (make_mn3_tmap): the producer
                        //    thread issues one copy per k-stage instead of bn / 32
  int cpt;              // 32-channel chunks per filter tap (TMA_A_IM2COL_K: k-stage = tap * cpt + chunk)
  int out_mode;         // 3: transposed convolution output: m = co, n = flat (image, pixel); bias is per m
                        // 1: backward-filter through TMA: m = (tap * cpt + chunk) * 32 + channel-in-chunk
                        // 2: the same over a space-to-depth view (below): virtual (tap, channel) -> real filter element
  int r_ci, r_fh, r_fw, r_sv, r_sh;   // out_mode 2: the real convolution's channels, filter and strides
  int pair_remote;      // CTA pair, tuning: 1 = the peer's copies count their bytes on the leader's barrier (cta_group::2 copy forms)
  int a_g3;             // 1 (TMA_A_TILED_MN): the tile's 4 (tall: 8) slabs arrive as one box of a make_mn3_tmap view
  int b_im2col;         // 1 (all-TMA path): B is the im2col operand -- bn output pixels x 32 channels of one tap per k-stage -- and A the
                        // packed filter (TMA_A_TILED_K): the TRANSPOSED orientation D[co][pixel] for narrow outputs (out_mode 3)
  int b_flip;           // B_FILTER_T: 1 = taps rotated by 180 degrees (stride-1 backward-data run as a forward convolution)
};

constexpr int BM = 128;        // UMMA M (cta_group::1)
constexpr int BK = 32;         // floats per stage row = 128 bytes = one swizzle span
constexpr int BN_MAX = 256;    // UMMA N limit
constexpr int kStages = 4;
constexpr int kABytes = BM * BK * 4;       // 16 KB
constexpr int kBBytes = BN_MAX * BK * 4;   // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kKtabMax = 8192;            // entries of the im2col k-decomposition table (32 KB)
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kKtabMax * 4;
constexpr int kThreads = 704;
constexpr int kProducerWarp0 = 6;
constexpr int kProducerThreads = 512;
constexpr int kTmemCols = 512;  // two accumulator buffers of 256 fp32 columns

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Producer-side arrive.  The default .release.cta form compiles to MEMBAR.ALL.CTA, which also waits for
// the look-ahead LDGs of later stages still in flight and serialises the gather pipeline.  What the
// consumer needs is only (1) this warp's st.shared performed before the arrive -- same thread, same
// shared-memory pipe, program order -- and (2) generic->async proxy visibility, given by
// fence.proxy.async.shared::cta right before.  Hence .relaxed.
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Time-bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t hint = 0x989680u) {
  // try_wait suspends the thread in hardware until the phase completes or the hint (ns) expires, so a
  // long hint keeps the poll loop out of the issue slots the gather warps need.
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(hint)
        : "memory");
    if (done) return;
    uint64_t now = globaltimer_ns();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 4000000000ull) asm volatile("trap;");  // 4 s: a protocol bug traps instead of hanging
  }
}
// the same wait with cluster-scope acquire: the phase is completed by an arrive from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, uint32_t hint = 0x989680u) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(hint)
        : "memory");
    if (done) return;
    uint64_t now = globaltimer_ns();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 4000000000ull) asm volatile("trap;");
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA pair (cta_group::2): two CTAs of a cluster on one TPC run ONE UMMA of 256 rows -- each holds its 128 rows of A and
// HALF of the B tile in its own shared memory and its 128 x N accumulator rows in its own TMEM; the leader (cluster rank 0)
// issues the instruction for both.  Per SM and flop the B bytes halve (the bound of the 128-row kernel), the ring stays
// 6 stages deep and both accumulator buffers remain, unlike the 256-row single-CTA tile.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the address, in the cluster's shared window, of the same shared-memory offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Fire-and-forget form: a release at cluster scope holds the thread for ~700 cycles per arrive.  Used where what the
// arrive publishes is already complete in hardware terms: bytes a TMA copy has landed in shared memory (observed through
// the local barrier) or TMEM loads retired by tcgen05.wait::ld.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// completion of the pair's MMAs arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, descriptors given as (low word, shared high word): the issuing thread only adds to the 14-bit address fields
__device__ __forceinline__ void umma_tf32_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// fp32 -> tf32 round-to-nearest (ties away from zero, == cvt.rna.tf32.f32) as ONE integer add of half a
// tf32 ulp: kind::tf32 reads only the top 19 bits of each operand word, so the low 13 bits need not be
// cleared.  inf stays inf (0x7f800000 + 0x1000 still has the inf pattern in its top 19 bits).
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout_type [61,64) with SWIZZLE_128B = 2).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;                 // LBO: unused for swizzled K-major
  d |= static_cast<uint64_t>(1024 >> 4) << 32;         // SBO
  d |= static_cast<uint64_t>(1) << 46;                 // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                 // SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2,
// B=TF32 [10,13)=2, both K-major, N>>3 at [17,23), M>>4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc(int n, bool a_mn_major = false, bool b_mn_major = false, int m = BM) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn_major ? 1u << 15 : 0u) | (b_mn_major ? 1u << 16 : 0u) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
// MN-major tf32 operand.  32-bit MN-major operands exist in one shared-memory layout only, SWIZZLE_128B_BASE32B
// (32-byte chunks swizzled inside the 128 B span by row & 3; TMA writes it as CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B):
// a row is 32 consecutive m of one k (128 B), 4 k rows make a 512 B swizzle atom (SBO = 512 B between atoms
// along k), the next 32-m chunk starts LBO = 4096 B later.  One tf32 UMMA (K = 8) reads two atoms per chunk.
__device__ __forceinline__ uint64_t make_sw128_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(4096 >> 4) << 16;         // LBO
  d |= static_cast<uint64_t>(512 >> 4) << 32;          // SBO
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;                 // SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of 16-byte chunk `kq` of row `r` inside a swizzled tile
__device__ __forceinline__ uint32_t sw128_off(int r, int kq) { return static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(kq ^ (r & 7)) << 4); }

// ------------------------------------------------------------------------------------------------
// operand element definitions (slow, fully decoded): used by the SIMT checker kernel and as the
// readable statement of what the fast gathers below compute.
// ------------------------------------------------------------------------------------------------
template <int AM>
__device__ __forceinline__ float a_elem(const GemmParams& p, int m, int k) {
  if (m >= p.M || k >= p.K) return 0.f;
  if (AM == A_COLMAJOR) return __ldg(p.a + m + static_cast<size_t>(k) * p.lda);
  const int ff = p.fh * p.fw;
  if (AM == A_IM2COL_FWD) {
    int img = m / p.P, pix = m - img * p.P, oh = pix / p.Wo, ow = pix - oh * p.Wo;
    int ci = k / ff, rs = k - ci * ff, r = rs / p.fw, s = rs - r * p.fw;
    int ih = oh * p.sv - p.ph + (p.fh - 1 - r), iw = ow * p.sh - p.pw + (p.fw - 1 - s);
    if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) return 0.f;
    return __ldg(p.a + ((static_cast<size_t>(img) * p.Ci + ci) * p.H + ih) * p.W + iw);
  }
  if (AM == A_IM2COL_BWD) {
    int img = m / p.P, pix = m - img * p.P, h = pix / p.W, w = pix - h * p.W;
    int co = k / ff, rs = k - co * ff, r = rs / p.fw, s = rs - r * p.fw;
    int t = h + p.ph - (p.fh - 1 - r), u = w + p.pw - (p.fw - 1 - s);
    if (t < 0 || u < 0 || t % p.sv || u % p.sh) return 0.f;
    int i = t / p.sv, j = u / p.sh;
    if (i >= p.Ho || j >= p.Wo) return 0.f;
    return __ldg(p.a + ((static_cast<size_t>(img) * p.Co + co) * p.Ho + i) * p.Wo + j);
  }
  {  // A_IM2COL_WGRAD: m = (ci,r,s), k = (img,oh,ow)
    int ci = m / ff, rs = m - ci * ff, r = rs / p.fw, s = rs - r * p.fw;
    int hw = p.Ho * p.Wo;
    int img = k / hw, pix = k - img * hw, oh = pix / p.Wo, ow = pix - oh * p.Wo;
    int ih = oh * p.sv - p.ph + (p.fh - 1 - r), iw = ow * p.sh - p.pw + (p.fw - 1 - s);
    if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) return 0.f;
    return __ldg(p.a + ((static_cast<size_t>(img) * p.Ci + ci) * p.H + ih) * p.W + iw);
  }
}
template <int BMD>
__device__ __forceinline__ float b_elem(const GemmParams& p, int n, int k) {
  if (n >= p.N || k >= p.K) return 0.f;
  if (BMD == B_KMAJOR) return __ldg(p.b + static_cast<size_t>(n) * p.ldb + k);
  if (BMD == B_FILTER_T) {
    const int ff = p.fh * p.fw, co = k / ff, rs = k - co * ff;
    return __ldg(p.b + (static_cast<size_t>(co) * p.N + n) * ff + (p.b_flip ? ff - 1 - rs : rs));
  }
  int hw = p.Ho * p.Wo;
  int img = k / hw, pix = k - img * hw;
  return __ldg(p.b + (static_cast<size_t>(img) * p.Co + n) * hw + pix);
}
__device__ __forceinline__ bool out_row_ok(const GemmParams& p, int m) {
  if (m >= p.M) return false;
  if (p.out_mode == 2) {   // virtual channel (c, dy, dx) of virtual tap (a, b) is real tap (a*sv + dy, b*sh + dx) of channel c
    const int chunk = m >> 5, tap = chunk / p.cpt, ch = (chunk - tap * p.cpt) * 32 + (m & 31);
    const int a = tap / p.fw, b = tap - a * p.fw, r = ch / p.r_sh;   // r = c * sv + dy
    return ch < p.Ci && a * p.r_sv + r % p.r_sv < p.r_fh && b * p.r_sh + (ch - r * p.r_sh) < p.r_fw;
  }
  return p.out_mode == 0 || p.out_mode == 3 || ((m >> 5) % p.cpt) * 32 + (m & 31) < p.Ci;   // padded channels of a tap carry no output
}
__device__ __forceinline__ size_t out_index(const GemmParams& p, int m, int n) {
  if (p.out_mode == 1) {   // filter_diff[co = n][ci][r][s]; tap (kh,kw) of the correlation is filter element ff-1-tap
    int chunk = m >> 5, tap = chunk / p.cpt, ci = (chunk - tap * p.cpt) * 32 + (m & 31), ff = p.fh * p.fw;
    return static_cast<size_t>(n) * p.col_stride + static_cast<size_t>(ci) * ff + (ff - 1 - tap);
  }
  if (p.out_mode == 2) {
    const int chunk = m >> 5, tap = chunk / p.cpt, ch = (chunk - tap * p.cpt) * 32 + (m & 31);
    const int a = tap / p.fw, b = tap - a * p.fw, r = ch / p.r_sh, c = r / p.r_sv;
    const int kh = a * p.r_sv + (r - c * p.r_sv), kw = b * p.r_sh + (ch - r * p.r_sh), ff = p.r_fh * p.r_fw;
    return static_cast<size_t>(n) * p.col_stride + static_cast<size_t>(c) * ff + (ff - 1 - (kh * p.r_fw + kw));
  }
  int img = m / p.P, pix = m - img * p.P;
  return static_cast<size_t>(img) * p.img_stride + pix + static_cast<size_t>(n) * p.col_stride;
}

// SIMT checker: one thread per output, sequential fp32 k loop.  Debug / cross-check only
// (mnv_debug_set_option("simt", 1)); exists in the tuning build only.
#ifdef MNV_TUNING
template <int AM, int BMD>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const GemmParams p) {
  size_t total = static_cast<size_t>(p.M) * p.N;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int m = static_cast<int>(t % p.M), n = static_cast<int>(t / p.M);
    float acc = 0.f;
    for (int k = 0; k < p.K; ++k) acc = fmaf(a_elem<AM>(p, m, k), b_elem<BMD>(p, n, k), acc);
    float v = acc + (p.bias ? __ldg(p.bias + n) : 0.f);
    p.out[out_index(p, m, n)] = (p.relu && !(v > 0.f)) ? 0.f : v;
  }
}
#endif

// ------------------------------------------------------------------------------------------------
// fast gathers.  Every producer thread stages 16 floats of A per k-stage (4 chunks of 16 bytes):
//   MC mapping (A_COLMAJOR, A_IM2COL_FWD, A_IM2COL_BWD): thread = one tile row, 16 consecutive k;
//      lanes run along m, which is the contiguous axis of the source => coalesced LDG.
//   KC mapping (A_IM2COL_WGRAD): thread = chunk column kq of 4 rows; 8 consecutive lanes cover 32
//      consecutive pixels (k), the contiguous axis of x for this mode.
// B (when it is not fetched by TMA): thread = 16-byte chunk (row, kq), rows b_row0 + 32*i.
// ------------------------------------------------------------------------------------------------
struct ARow {           // per-thread, per-tile state of the A gather
  bool valid;
  long long base;       // element offset contribution of the row
  uint32_t mh, mw;      // validity masks over kh / kw (im2col fwd / bwd)
  int h0, w0;           // oh*sv-ph (fwd), h+ph (bwd)
};
template <int NR>
struct AWgrad {         // KC mapping: NR rows (taps) per thread
  int toff[NR];         // ci*H*W + (kh-ph)*W + (kw-pw); row past M: a bound-failing dh instead of a flag
  int dhw[NR];          // (kh-ph) << 16 | ((kw-pw) & 0xffff)
};

template <int AM>
__device__ __forceinline__ ARow a_row_setup(const GemmParams& p, int m) {
  ARow r;
  r.valid = m < p.M;
  r.base = 0; r.mh = r.mw = 0; r.h0 = r.w0 = 0;
  if (!r.valid) return r;
  if (AM == A_COLMAJOR) {
    r.base = m;
  } else if (AM == A_IM2COL_FWD) {
    int img = m / p.P, pix = m - img * p.P, oh = pix / p.Wo, ow = pix - oh * p.Wo;
    r.h0 = oh * p.sv - p.ph; r.w0 = ow * p.sh - p.pw;
    r.base = static_cast<long long>(img) * p.Ci * p.H * p.W + static_cast<long long>(r.h0) * p.W + r.w0;
    for (int kh = 0; kh < p.fh; ++kh) if (r.h0 + kh >= 0 && r.h0 + kh < p.H) r.mh |= 1u << kh;
    for (int kw = 0; kw < p.fw; ++kw) if (r.w0 + kw >= 0 && r.w0 + kw < p.W) r.mw |= 1u << kw;
  } else if (AM == A_IM2COL_BWD) {
    int img = m / p.P, pix = m - img * p.P, h = pix / p.W, w = pix - h * p.W;
    r.h0 = h + p.ph; r.w0 = w + p.pw;
    r.base = static_cast<long long>(img) * p.Co * p.Ho * p.Wo;
    for (int kh = 0; kh < p.fh; ++kh) { int t = r.h0 - kh; if (t >= 0 && t % p.sv == 0 && t / p.sv < p.Ho) r.mh |= 1u << kh; }
    for (int kw = 0; kw < p.fw; ++kw) { int u = r.w0 - kw; if (u >= 0 && u % p.sh == 0 && u / p.sh < p.Wo) r.mw |= 1u << kw; }
  }
  return r;
}

template <int NR>
__device__ __forceinline__ AWgrad<NR> a_wgrad_setup(const GemmParams& p, int m_base, int row0, int row_step) {
  AWgrad<NR> w;
  const int ff = p.fh * p.fw;
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    int m = m_base + row0 + row_step * i;
    bool valid = m < p.M;
    int mm = valid ? m : 0;
    int ci = mm / ff, rs = mm - ci * ff, rr = rs / p.fw, ss = rs - rr * p.fw;
    int dh = valid ? (p.fh - 1 - rr) - p.ph : -20000;   // invalid row: every bounds test fails
    int dw = (p.fw - 1 - ss) - p.pw;
    w.dhw[i] = (dh << 16) | (dw & 0xffff);
    w.toff[i] = valid ? ci * p.H * p.W + dh * p.W + dw : 0;
  }
  return w;
}

// MC mapping: NE consecutive k starting at k0 for this thread's row
template <int AM, int NE>
__device__ __forceinline__ void a_gatherN(const GemmParams& p, const ARow& r, int k0, float (&v)[NE]) {
  if (AM == A_COLMAJOR) {
    const float* src = p.a + r.base + static_cast<size_t>(k0) * p.lda;
#pragma unroll
    for (int j = 0; j < NE; ++j) v[j] = (r.valid && k0 + j < p.K) ? __ldg(src + static_cast<size_t>(j) * p.lda) : 0.f;
  } else {
    // k = (c, rr, ss) in filter storage order; kh = fh-1-rr, kw = fw-1-ss
    const int ff = p.fh * p.fw;
    int c = k0 / ff, rs = k0 - c * ff, rr = rs / p.fw, ss = rs - rr * p.fw;
    int kh = p.fh - 1 - rr, kw = p.fw - 1 - ss;
    if (AM == A_IM2COL_FWD) {
      const int HW = p.H * p.W;
      int off = c * HW + kh * p.W + kw;
      const float* src = p.a + r.base;
#pragma unroll
      for (int j = 0; j < NE; ++j) {  // table-free fallback (K or C*H*W too large for the smem table)
        bool ok = r.valid && (k0 + j < p.K) && (((r.mh >> kh) & (r.mw >> kw)) & 1u);
        v[j] = ok ? __ldg(src + off) : 0.f;
        --kw; --off;
        if (kw < 0) { kw = p.fw - 1; off += p.fw - p.W; --kh; if (kh < 0) { kh = p.fh - 1; off += p.fh * p.W + HW; } }
      }
    } else {  // strided backward-data: divisions per element, first-layer shapes only
      const int HoWo = p.Ho * p.Wo;
      const float* src = p.a + r.base;
#pragma unroll
      for (int j = 0; j < NE; ++j) {
        bool ok = r.valid && (k0 + j < p.K) && (((r.mh >> kh) & (r.mw >> kw)) & 1u);
        float val = 0.f;
        if (ok) {
          int i = (r.h0 - kh) / p.sv, jj = (r.w0 - kw) / p.sh;
          val = __ldg(src + static_cast<size_t>(c) * HoWo + i * p.Wo + jj);
        }
        v[j] = val;
        --kw;
        if (kw < 0) { kw = p.fw - 1; --kh; if (kh < 0) { kh = p.fh - 1; ++c; } }
      }
    }
  }
}

// Forward im2col with the k-decomposition table in shared memory (built once per CTA), one word per k:
//   KT == 1 (fh*fw <= 31): (c*H*W + kh*W + kw) << 5 | (kh*fw + kw); the tap index is tested against ONE
//            combined per-row mask (bit kh*fw+kw = tap in bounds);
//   KT == 2: (c*H*W + kh*W + kw) | kh << 22 | kw << 27, tested against the row's kh and kw masks.
//   Entries past K select bit 31, which no mask has.
// Row validity is folded into the masks (zero for rows past M), so one element costs two shifts, a
// predicate, a 64-bit shift-add and a predicated LDG; the table words are one LDS.128 per four elements.
template <int KT, int NE>
__device__ __forceinline__ void a_gatherN_ktab(const float* __restrict__ src, uint32_t m0, uint32_t m1, uint32_t ktab_addr,
                                               float (&v)[NE]) {
#pragma unroll
  for (int q = 0; q < NE / 4; ++q) {
    uint4 e4 = ld_shared_v4(ktab_addr + 16 * q);
    const uint32_t e[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bool ok;
      uint32_t off;
      if (KT == 1) {
        ok = __funnelshift_r(m0, 0u, e[j]) & 1u;     // shift amount = low 5 bits = tap index
        off = e[j] >> 5;
      } else {
        ok = (__funnelshift_r(m0, 0u, e[j] >> 22) & (m1 >> (e[j] >> 27))) & 1u;
        off = e[j] & 0x3FFFFFu;
      }
      v[4 * q + j] = ok ? __ldg(src + off) : 0.f;
    }
  }
}

// KC mapping (backward-filter): this thread's 4 pixels for its kARows tap rows; v[4*i + e].
// spi == 0: k = (img, oh, ow) flat, k0 = ks*32 + kq*4.  spi > 0 (top_diff comes by TMA): every image's
// pixels are padded to spi*32, so a k-stage never straddles two images.
template <int NR>
__device__ __forceinline__ void a_gather_wgrad(const GemmParams& p, const AWgrad<NR>& w, int ks, int kq, float (&v)[4 * NR]) {
  const int HoWo = p.Ho * p.Wo;
  int img, pix;
  bool in_k;
  if (p.spi > 0) {
    img = ks / p.spi;
    pix = (ks - img * p.spi) * BK + kq * 4;
    in_k = true;
  } else {
    int k0 = ks * BK + kq * 4;
    img = k0 / HoWo;
    pix = k0 - img * HoWo;
    in_k = false;
  }
  int oh = pix / p.Wo, ow = pix - oh * p.Wo;
  int uoff[4], ohs[4], ows[4];
  const float* src = p.a + static_cast<long long>(img) * p.Ci * p.H * p.W;   // image base (k-stage never straddles when spi > 0)
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    bool ok = in_k ? (pix + e < HoWo) : (ks * BK + kq * 4 + e < p.K);
    ohs[e] = ok ? oh * p.sv : -(1 << 28);   // out-of-range k fails the bounds test below
    ows[e] = ow * p.sh;
    uoff[e] = oh * p.sv * p.W + ow * p.sh;
    if (++ow == p.Wo) {
      ow = 0;
      if (++oh == p.Ho) { oh = 0; uoff[e] += 0; if (!in_k) { /* next image */ } }
    }
  }
  // unpadded K (spi == 0): pixels past the end of the image belong to the next one
  if (!in_k) {
    int k0 = ks * BK + kq * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int k = k0 + e, im = k / HoWo, px = k - im * HoWo, o2 = px / p.Wo, w2 = px - o2 * p.Wo;
      uoff[e] = (im - img) * p.Ci * p.H * p.W + o2 * p.sv * p.W + w2 * p.sh;
      ohs[e] = (k < p.K) ? o2 * p.sv : -(1 << 28);
      ows[e] = w2 * p.sh;
    }
  }
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const float* s2 = src + w.toff[i];
    const int dh = w.dhw[i] >> 16, dw = static_cast<int>(static_cast<short>(w.dhw[i] & 0xffff));
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      bool ok = static_cast<unsigned>(ohs[e] + dh) < static_cast<unsigned>(p.H) &&
                static_cast<unsigned>(ows[e] + dw) < static_cast<unsigned>(p.W);
      v[4 * i + e] = ok ? __ldg(s2 + uoff[e]) : 0.f;
    }
  }
}

// B gathered by threads: `iters` chunks (rows b_row0 + 32*i, column chunk kq) of 4 consecutive k
template <int BMD>
__device__ __forceinline__ void b_gather(const GemmParams& p, int n_base, int b_row0, int iters, int k0, float4 (&vb)[4]) {
  if (BMD == B_KMAJOR) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i >= iters) break;
      int row = b_row0 + 64 * i, n = n_base + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < p.bn && n < p.N && k0 < p.K) {
        const float* src = p.b + static_cast<size_t>(n) * p.ldb + k0;
        if (p.b_vec) {
          v = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          v.x = __ldg(src);
          if (k0 + 1 < p.K) v.y = __ldg(src + 1);
          if (k0 + 2 < p.K) v.z = __ldg(src + 2);
          if (k0 + 3 < p.K) v.w = __ldg(src + 3);
        }
      }
      vb[i] = v;
    }
  } else if (BMD == B_FILTER_T) {  // k = (co, r, s): filter[(co*N + n)*ff + tap]; same k decomposition for all rows
    const int ff = p.fh * p.fw;
    int co = k0 / ff, rs = k0 - co * ff;
    long long off[4];
    bool okk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      okk[e] = k0 + e < p.K;
      off[e] = static_cast<long long>(co) * p.N * ff + (p.b_flip ? ff - 1 - rs : rs);
      if (++rs == ff) { rs = 0; ++co; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i >= iters) break;
      int row = b_row0 + 64 * i, n = n_base + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < p.bn && n < p.N) {
        const float* src = p.b + static_cast<size_t>(n) * ff;
        if (okk[0]) v.x = __ldg(src + off[0]);
        if (okk[1]) v.y = __ldg(src + off[1]);
        if (okk[2]) v.z = __ldg(src + off[2]);
        if (okk[3]) v.w = __ldg(src + off[3]);
      }
      vb[i] = v;
    }
  } else {  // B_DY_WGRAD: k = (img, pix): top_diff[(img*Co + n)*hw + pix]; same k for all of this thread's rows
    const int hw = p.Ho * p.Wo;
    int img = k0 / hw, pix = k0 - img * hw;
    long long off[4];
    bool okk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      okk[e] = k0 + e < p.K;
      off[e] = static_cast<long long>(img) * p.Co * hw + pix;
      if (++pix == hw) { pix = 0; ++img; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i >= iters) break;
      int row = b_row0 + 64 * i, n = n_base + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < p.bn && n < p.N) {
        const float* src = p.b + static_cast<size_t>(n) * hw;
        if (okk[0]) v.x = __ldg(src + off[0]);
        if (okk[1]) v.y = __ldg(src + off[1]);
        if (okk[2]) v.z = __ldg(src + off[2]);
        if (okk[3]) v.w = __ldg(src + off[3]);
      }
      vb[i] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TMA (B operand when it is a plain K-major matrix with 16-byte-aligned rows)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// CTA-pair forms: the copy lands in the executing CTA's shared memory, its bytes are counted on the LEADER's barrier
// (`bar` = that barrier's address in the cluster window, mapa_u32(bar, 0)).
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_2sm(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c, int w, int h,
                                                       int n, int kw, int kh) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n),
        "h"(static_cast<uint16_t>(kw)), "h"(static_cast<uint16_t>(kh))
      : "memory");
}

// L2 prefetch of a box a few k-stages ahead: the smem ring covers ~1.5 us of TMA latency, less than an HBM round trip
// under load, so operands that stream from HBM (FC weights, channels-last activation copies) are pulled into L2 early.
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tmap, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_im2col_4d(const CUtensorMap* tmap, int c, int w, int h, int n, int kw, int kh) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.im2col [%0, {%1, %2, %3, %4}], {%5, %6};"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(static_cast<uint16_t>(kw)), "h"(static_cast<uint16_t>(kh))
               : "memory");
}

// im2col-mode load: the box starts at pixel (w, h, n) of the (C, W, H, N) tensor -- a position of the filter
// window's origin inside the bounding box the map was encoded with -- shifted by the tap offsets (kw, kh), and
// walks `pixelsPerColumn` window positions along W, then H, then N; out-of-tensor elements arrive as zeros.
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c, int w, int h,
                                                   int n, int kw, int kh) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n),
        "h"(static_cast<uint16_t>(kw)), "h"(static_cast<uint16_t>(kh))
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
constexpr int kRasterM = 16;
struct TileCoord { int mt, nt, split, ks_begin, ks_end, partial; };
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile) {
  TileCoord t;
  // Grouped rasterisation: inside one K split the tile index walks kRasterM m-tiles down, then one n-tile
  // across, so a wave of 148 CTAs covers a ~16 x 9 block of tiles instead of 4.6 full rows: the operand rows
  // a wave touches drop from (4.6*128 + N) to (16*128 + 9*bn), which is what keeps an 8192^3 GEMM's B
  // (268 MB > L2) from being re-fetched from HBM every wave.
  const int per_split = p.m_tiles * p.n_tiles;
  int id, stages;
  if (tile >= p.tail_first) {          // tail split (implies splits == 1): the last tiles, tail_splits CTAs each
    const int u = tile - p.tail_first, ntail = per_split - p.tail_first;
    t.split = u / ntail;
    id = p.tail_first + (u - t.split * ntail);
    stages = p.tail_stages;
    t.partial = 1;
  } else {
    t.split = tile / per_split;
    id = tile - t.split * per_split;
    stages = p.stages_per_split;
    t.partial = p.splits > 1;
  }
  const int band = kRasterM * p.n_tiles;
  const int g = id / band;
  id -= g * band;
  const int first_m = g * kRasterM;
  const int gm = min(kRasterM, p.m_tiles - first_m);
  t.nt = id / gm;
  t.mt = first_m + (id - t.nt * gm);
  t.ks_begin = t.split * stages;
  t.ks_end = min(p.k_stages, t.ks_begin + stages);
  return t;
}

template <int AM, int BMD, bool BTMA, int RING>
__global__ void __launch_bounds__(kThreads, 1) umma_gemm_kernel(const __grid_constant__ GemmParams p,
                                                                const __grid_constant__ CUtensorMap tmap_b,
                                                                const __grid_constant__ CUtensorMap tmap_a) {
  pdl_trigger();
  // WIDE: one 128 x bn tile (256 < bn <= 384) as two UMMA halves sharing the A tile, single TMEM accumulator,
  // 3-deep ring of 64 KB slots; otherwise bn <= 256, two accumulators, 4-deep ring of 48 KB slots.
  // RING 2 (both operands through TMA, bn <= 128): 6-deep ring of 32 KB slots -- a 128 x 96 tile consumes a stage
  // in ~200 cycles, so four stages in flight do not cover the TMA round trip.
  // RING 3 (both operands through TMA): a 256 x bn tile as two UMMA halves of 128 rows sharing the B tile, one
  // accumulator per half (all 512 TMEM columns), 3-deep ring of 64 KB slots.  The kernel is bound by L2 -> SM
  // operand traffic (48 KB per 128x256x32 stage = 44 flop/B); the tall tile moves 64 KB for twice the math.
  // RING 4 (both operands through TMA): the CTA PAIR -- a 256 x bn tile computed by two CTAs of a cluster with
  // tcgen05.mma.cta_group::2: this CTA stages its 128 rows of A and bn / 2 rows of B (32 KB per stage, 6-deep ring) and
  // drains its own 128 accumulator rows; rank 0 issues the MMAs for both and owns the full / tmem_empty barriers.
  constexpr bool WIDE = RING == 1, DEEP = RING == 2, TALL = RING == 3, PAIR = RING == 4;
  static_assert(!PAIR || AM == A_TMA, "the CTA pair exists on the all-TMA path only");
  constexpr int kTileM = (TALL || PAIR) ? 2 * BM : BM;
  constexpr int kATile = TALL ? 2 * kABytes : kABytes;
  constexpr int kNStages = (WIDE || TALL) ? 3 : (DEEP || PAIR) ? 6 : kStages;
  constexpr int kSBytes = WIDE ? kABytes + 384 * BK * 4 : (DEEP || PAIR) ? kABytes + 128 * BK * 4 : TALL ? 2 * kABytes + kBBytes : kStageBytes;
  constexpr int kNAcc = (WIDE || TALL) ? 1 : 2;
  constexpr int kMaxStages = 6;
  // With both operands on TMA the 16 gather warps have no mainloop work: they join the epilogue, five warps per
  // TMEM lane quarter (a warp reaches lanes 32 * (warp % 4) .. +31 only), each draining every fifth 16-column chunk.
  constexpr int kEpiSlots = AM == A_TMA ? 5 : 1;
  constexpr int kEpiWarps = 4 * kEpiSlots;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: the 128B swizzle pattern is a function of address bits [7,10)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  // bars[0..3] full, [4..7] empty, [8..9] tmem_full, [10..11] tmem_empty, then the TMEM base slot
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kMaxStages);
  const uint32_t tfull0 = smem_u32(bars + 2 * kMaxStages), tempty0 = smem_u32(bars + 2 * kMaxStages + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  const uint32_t pfull0 = smem_u32(bars + 2 * kMaxStages + 6);   // CTA pair, leader: "the peer's half of stage s has landed"
  const uint32_t ktab0 = smem_u32(smem + kStages * kStageBytes + 256);   // uint32 ktab[kKtabMax]
  const uint32_t smem_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.total_units;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;                 // 0: the pair's leader
  // (expressions, not variables: the single-CTA instantiations keep reading blockIdx / gridDim / p.bn in place)
#define tile_first (PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x))
#define tile_step (PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x))
#define bn_cta (PAIR ? p.bn >> 1 : p.bn)                               /* B rows staged by this CTA */
  const int m_rank = PAIR ? static_cast<int>(rank) * BM : 0;           // this CTA's rows inside the pair's tile
  const int n_rank = PAIR ? static_cast<int>(rank) * bn_cta : 0;

  if (threadIdx.x == 0) {
    // full: TMA-fed B: the 4 warps of the slot's producer group + the TMA thread's arrive.expect_tx;
    //       gathered B: all 16 producer warps
    for (int s = 0; s < kNStages; ++s) { mbar_init(full0 + 8 * s, AM == A_TMA ? 1 : BTMA ? 4 + 1 : kProducerThreads / 32); mbar_init(empty0 + 8 * s, 1); }   // kNStages of them are used
    for (int s = 0; s < 2; ++s) { mbar_init(tfull0 + 8 * s, 1); mbar_init(tempty0 + 8 * s, PAIR ? 2 * kEpiWarps : kEpiWarps); }
    if (PAIR) for (int s = 0; s < kNStages; ++s) mbar_init(pfull0 + 8 * s, 1);
    fence_barrier_init();
  }
  if (warp == 4) { if (PAIR) tmem_alloc_2sm(smem_u32(tmem_slot), kTmemCols); else tmem_alloc(smem_u32(tmem_slot), kTmemCols); }
  if ((AM == A_IM2COL_FWD || AM == A_IM2COL_FWD_K) && p.use_ktab) {
    const int ff = p.fh * p.fw, HW = p.H * p.W;
    for (int k = threadIdx.x; k < p.k_stages * BK; k += blockDim.x) {
      uint32_t e = p.use_ktab == 1 ? 31u : (31u << 22);
      if (k < p.K) {
        int c = k / ff, rs = k - c * ff, rr = rs / p.fw, ss = rs - rr * p.fw;
        int kh = p.fh - 1 - rr, kw = p.fw - 1 - ss;
        uint32_t off = static_cast<uint32_t>(c * HW + kh * p.W + kw);
        e = p.use_ktab == 1 ? (off << 5 | static_cast<uint32_t>(kh * p.fw + kw))
                            : (off | static_cast<uint32_t>(kh) << 22 | static_cast<uint32_t>(kw) << 27);
      }
      st_shared_u32(ktab0 + 4 * k, e);
    }
  }
  pdl_wait();          // everything above touched only shared memory / TMEM / kernel parameters
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ===================== epilogue (run by warps 0-3, and by the idle producer warps on the all-TMA path) ==========
  auto epilogue = [&](const int quarter, const int slot) {
    int acc_stage = 0; uint32_t acc_phase = 0;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      TileCoord t = decode_tile(p, tile);
      mbar_wait(tfull0 + 8 * acc_stage, acc_phase, p.wait_hint);
      tc_fence_after();
      const int n0 = t.nt * p.bn;
      const bool add_bias = p.bias != nullptr && !t.partial;
      const bool relu = p.relu != 0 && !t.partial;
#pragma unroll
      for (int half = 0; half < (TALL ? 2 : 1); ++half) {
        const int m = t.mt * kTileM + m_rank + half * BM + quarter * 32 + lane;
        const bool row_ok = out_row_ok(p, m);
        float* dst;
        long long cstride;
        if (t.partial) {  // partial[split][n][m]
          dst = p.partial + static_cast<size_t>(t.split) * p.M * p.N + (row_ok ? m : 0);
          cstride = p.M;
        } else {
          dst = p.out + (row_ok ? out_index(p, m, 0) : 0);
          cstride = p.col_stride;
        }
        for (int c0 = 16 * slot; c0 < p.bn; c0 += 16 * kEpiSlots) {
          uint32_t r[16];
          tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc_stage * BN_MAX + half * BN_MAX + c0), r);
          if (p.out_mode == 3) {   // lane = output channel m, columns = 16 consecutive flat pixels: contiguous along the pixel axis
            if (row_ok) {          // of plane (image, m) until the image ends (never split-K: see conv_tma_fprop)
              const float bv = add_bias ? __ldg(p.bias + m) : 0.f;
              int n = n0 + c0, img = n / p.P, pix = n - img * p.P;
              float* q = p.out + static_cast<size_t>(img) * p.img_stride + static_cast<size_t>(m) * p.P;
              float v[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float val = __uint_as_float(r[j]) + bv;
                v[j] = (relu && !(val > 0.f)) ? 0.f : val;
              }
              if (pix + 16 <= p.P && n + 16 <= p.N) {
                // the 16 pixels are one contiguous 64-byte run of plane (img, m): 16-byte stores from the first aligned
                // float on (the run's alignment differs per lane: plane sizes are odd), scalars at both ends
                float* d = q + pix;
                const int head = static_cast<int>((4u - ((reinterpret_cast<uintptr_t>(d) >> 2) & 3u)) & 3u);
#define MNV_RUN(H)                                                                                                         \
                {                                                                                                          \
                  _Pragma("unroll") for (int j = 0; j < H; ++j) d[j] = v[j];                                               \
                  _Pragma("unroll") for (int j = H; j + 4 <= 16; j += 4)                                                   \
                    *reinterpret_cast<float4*>(d + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);                   \
                  _Pragma("unroll") for (int j = H + (16 - H) / 4 * 4; j < 16; ++j) d[j] = v[j];                           \
                }
                if (head == 0) MNV_RUN(0) else if (head == 1) MNV_RUN(1) else if (head == 2) MNV_RUN(2) else MNV_RUN(3)
#undef MNV_RUN
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j, ++n) {
                  if (n < p.N) {
                    q[pix] = v[j];
                    if (++pix == p.P) { pix = 0; q += p.img_stride; }
                  }
                }
              }
            }
            continue;
          }
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              int n = n0 + c0 + j;
              if (n < p.N) {
                float val = __uint_as_float(r[j]);
                if (add_bias) val += __ldg(p.bias + n);
                if (relu) val = val > 0.f ? val : 0.f;
                dst[static_cast<size_t>(n) * cstride] = val;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster_relaxed(mapa_u32(tempty0 + 8 * acc_stage, 0));    // the leader's MMA thread waits for both CTAs' drains
        else mbar_arrive(tempty0 + 8 * acc_stage);
      }
      if (++acc_stage == kNAcc) { acc_stage = 0; acc_phase ^= 1; }
    }
  };
  if (warp < 4) {
    epilogue(warp, 0);
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    // The whole warp walks the pipeline (so every lane reaches the final __syncthreads together);
    // lane 0 alone issues tcgen05.mma / tcgen05.commit.
    const int bnh = WIDE ? p.bn / 2 : p.bn;           // columns per UMMA instruction
    const bool a_mn = AM == A_TMA && (p.a_mode == TMA_A_TILED_MN || p.a_mode == TMA_A_IM2COL_MN);
    const bool b_mn = AM == A_TMA && p.b_mn != 0;
    const uint32_t idesc = make_idesc(bnh, a_mn, b_mn, PAIR ? 2 * BM : BM);
    int stage = 0; uint32_t phase = 0;
    int acc_stage = 0; uint32_t acc_phase = 0;
    for (int tile = tile_first; (!PAIR || rank == 0) && tile < total_tiles; tile += tile_step) {
      TileCoord t = decode_tile(p, tile);
      mbar_wait(tempty0 + 8 * acc_stage, acc_phase ^ 1, p.wait_hint);  // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc_stage * BN_MAX);
      for (int ks = t.ks_begin; ks < t.ks_end; ++ks) {
        mbar_wait(full0 + 8 * stage, phase, p.wait_hint);
        if (PAIR && !p.pair_remote) mbar_wait_cluster(pfull0 + 8 * stage, phase, p.wait_hint);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_base + stage * kSBytes;
          const uint32_t b_addr = a_addr + kATile;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {  // UMMA K = 8 tf32 = 32 bytes inside the swizzle span
            const uint32_t acc = (ks > t.ks_begin || kk > 0) ? 1u : 0u;
            const uint64_t adesc = a_mn ? make_sw128_desc_mn(a_addr + kk * 1024) : make_sw128_desc(a_addr + kk * 32);
            const uint64_t bdesc = b_mn ? make_sw128_desc_mn(b_addr + kk * 1024) : make_sw128_desc(b_addr + kk * 32);
            if (PAIR) umma_tf32_2sm(tmem_d, adesc, bdesc, idesc, acc);
            else umma_tf32(tmem_d, adesc, bdesc, idesc, acc);
            if (WIDE)   // second half of the columns: same A tile, B rows bnh.., TMEM columns bnh..
              umma_tf32(tmem_d + bnh, adesc, b_mn ? make_sw128_desc_mn(b_addr + bnh * 128 + kk * 1024) : make_sw128_desc(b_addr + bnh * 128 + kk * 32), idesc, acc);
            if (TALL)   // rows 128..255: second A half, same B tile, second accumulator
              umma_tf32(tmem_d + BN_MAX, a_mn ? make_sw128_desc_mn(a_addr + kABytes + kk * 1024) : make_sw128_desc(a_addr + kABytes + kk * 32),
                        bdesc, idesc, acc);
          }
          if (PAIR) {
            umma_commit_2sm(empty0 + 8 * stage);                              // both CTAs' slots, both CTAs' accumulators
            if (ks + 1 == t.ks_end) umma_commit_2sm(tfull0 + 8 * acc_stage);
          } else {
          umma_commit(empty0 + 8 * stage);  // frees the smem slot when these MMAs have read it
          if (ks + 1 == t.ks_end) umma_commit(tfull0 + 8 * acc_stage);  // accumulator complete
          }
        }
        __syncwarp();
        if (++stage == kNStages) { stage = 0; phase ^= 1; }
      }
      if (++acc_stage == kNAcc) { acc_stage = 0; acc_phase ^= 1; }
    }
    if (PAIR && rank != 0 && lane == 0 && !p.pair_remote) {
      // The peer's half of the pipeline hand-over: every copy counts its bytes on the barrier of the CTA it lands in
      // (copies that signal a barrier in the other CTA measured ~4x the per-copy cost); this thread forwards "stage
      // landed" to the leader, whose MMA thread waits for both halves.
      const uint32_t pfull_ld0 = mapa_u32(pfull0, 0);
      int st = 0; uint32_t ph = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        TileCoord t = decode_tile(p, tile);
        for (int ks = t.ks_begin; ks < t.ks_end; ++ks) {
          mbar_wait(full0 + 8 * st, ph, p.wait_hint);
          mbar_arrive_cluster_relaxed(pfull_ld0 + 8 * st);
          if (++st == kNStages) { st = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===================== TMA producer for B =====================
    if (BTMA && lane == 0) {
      int stage = 0; uint32_t phase = 0;
      // the whole box(es), OOB rows/cols arrive as zeros
#ifdef MNV_TUNING
      // diagnostic (tuning build only): pf_dist = -1 / -2 / -3 skips the A / B / both copies -- results are garbage, the time
      // says which operand stream bounds the mainloop
      const int dbg_skip = p.pf_dist < 0 ? -p.pf_dist : 0;
#else
      constexpr int dbg_skip = 0;
#endif
      const uint32_t bytes = ((dbg_skip & 2) ? 0u : static_cast<uint32_t>(bn_cta) * BK * 4) + ((AM == A_TMA && !(dbg_skip & 1)) ? kATile : 0);
      // CTA pair: where this CTA's copies count their bytes -- the leader's barrier (cta_group::2 copy forms) or its own
      const bool remote_bar = PAIR && p.pair_remote;
      const uint32_t full_ld0 = remote_bar ? mapa_u32(full0, 0) : full0;
#define MNV_LD2(...) do { if constexpr (PAIR) { if (remote_bar) tma_load_2d_2sm(__VA_ARGS__); else tma_load_2d(__VA_ARGS__); } else tma_load_2d(__VA_ARGS__); } while (0)
#define MNV_LD3(...) do { if constexpr (PAIR) { if (remote_bar) tma_load_3d_2sm(__VA_ARGS__); else tma_load_3d(__VA_ARGS__); } else tma_load_3d(__VA_ARGS__); } while (0)
#define MNV_LDI(...) do { if constexpr (PAIR) { if (remote_bar) tma_load_im2col_4d_2sm(__VA_ARGS__); else tma_load_im2col_4d(__VA_ARGS__); } else tma_load_im2col_4d(__VA_ARGS__); } while (0)
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        TileCoord t = decode_tile(p, tile);
        // A_TMA: position of the tile's first row
        constexpr int kHalves = TALL ? 2 : 1, kChunks = 4 * kHalves;
        int a_w[kHalves] = {}, a_h[kHalves] = {}, a_n[kHalves] = {};
        int a_kw[kChunks] = {}, a_kh[kChunks] = {}, a_c[kChunks] = {};
        if (AM == A_TMA) {
          if (p.a_mode == TMA_A_IM2COL_K) {
#pragma unroll
            for (int h = 0; h < kHalves; ++h) {   // rows past M land in image >= N: zero-filled
              const int m0 = t.mt * kTileM + m_rank + h * BM;
              a_n[h] = m0 / p.P;
              const int pix = m0 - a_n[h] * p.P, oh = pix / p.Wo;
              a_h[h] = oh * p.sv - p.ph; a_w[h] = (pix - oh * p.Wo) * p.sh - p.pw;
            }
          } else if (p.a_mode == TMA_A_IM2COL_MN) {
            const int last = p.fh * p.fw * p.cpt - 1;   // rows past M still receive a (masked) box
#pragma unroll
            for (int j = 0; j < kChunks; ++j) {
              const int chunk = min(t.mt * (PAIR ? 2 * kChunks : kChunks) + (m_rank >> 5) + j, last), tap = chunk / p.cpt;
              a_c[j] = (chunk - tap * p.cpt) * 32; a_kh[j] = tap / p.fw; a_kw[j] = tap - a_kh[j] * p.fw;
            }
          }
        }
        for (int ks = t.ks_begin; ks < t.ks_end; ++ks) {
          if (p.pf_dist > 0 && ks + p.pf_dist < t.ks_end) {   // pull the boxes of k-stage ks + pf_dist into L2
            const int kp = ks + p.pf_dist;
            if (AM == A_TMA) {
              if (p.a_mode == TMA_A_IM2COL_K) {
                const int tap = kp / p.cpt, cc = kp - tap * p.cpt, kh = tap / p.fw;
#pragma unroll
                for (int h = 0; h < kHalves; ++h) tma_prefetch_im2col_4d(&tmap_a, cc * BK, a_w[h], a_h[h], a_n[h], tap - kh * p.fw, kh);
              } else if (p.a_mode == TMA_A_TILED_MN) {
                if (p.a_g3) tma_prefetch_3d(&tmap_a, 0, kp * BK, (t.mt * kTileM + m_rank) >> 5);
                else
#pragma unroll
                for (int j = 0; j < kChunks; ++j) tma_prefetch_2d(&tmap_a, t.mt * kTileM + m_rank + 32 * j, kp * BK);
              } else if (p.a_mode == TMA_A_TILED_K) {
#pragma unroll
                for (int h = 0; h < kHalves; ++h) tma_prefetch_2d(&tmap_a, kp * BK, t.mt * kTileM + m_rank + h * BM);
              }
            }
            if (!(AM == A_TMA && (p.b_mn || p.b_im2col))) {
              const int halves = WIDE ? 2 : 1, rows = WIDE ? p.bn / 2 : p.bn;
              for (int h = 0; h < halves; ++h) {
                const int n0 = t.nt * p.bn + n_rank + h * rows;
                if (p.spi > 0) { const int img = kp / p.spi; tma_prefetch_3d(&tmap_b, (kp - img * p.spi) * BK, n0, img); }
                else tma_prefetch_2d(&tmap_b, kp * BK, n0);
              }
            }
          }
          mbar_wait(empty0 + 8 * stage, phase ^ 1, p.wait_hint);
          if (!remote_bar) mbar_arrive_expect_tx(full0 + 8 * stage, bytes);
          else if (rank == 0) mbar_arrive_expect_tx(full0 + 8 * stage, 2 * bytes);      // both CTAs' copies land on the leader's barrier
          const uint32_t fbar = full_ld0 + 8 * stage;
          if (AM == A_TMA && !(dbg_skip & 1)) {
            const uint32_t a_dst = smem_base + stage * kSBytes, bar = fbar;
            if (p.a_mode == TMA_A_IM2COL_K) {
              const int tap = ks / p.cpt, cc = ks - tap * p.cpt, kh = tap / p.fw;
#pragma unroll
              for (int h = 0; h < kHalves; ++h)
                MNV_LDI(a_dst + h * kABytes, &tmap_a, bar, cc * BK, a_w[h], a_h[h], a_n[h], tap - kh * p.fw, kh);
            } else if (p.a_mode == TMA_A_TILED_MN) {
              if (p.a_g3) MNV_LD3(a_dst, &tmap_a, bar, 0, ks * BK, (t.mt * kTileM + m_rank) >> 5);
              else
#pragma unroll
              for (int j = 0; j < kChunks; ++j) MNV_LD2(a_dst + j * 4096, &tmap_a, bar, t.mt * kTileM + m_rank + 32 * j, ks * BK);
            } else if (p.a_mode == TMA_A_TILED_K) {
#pragma unroll
              for (int h = 0; h < kHalves; ++h) MNV_LD2(a_dst + h * kABytes, &tmap_a, bar, ks * BK, t.mt * kTileM + m_rank + h * BM);
            } else {   // k-stage = 32 output pixels: of one image (p.spi stages per image, top_diff padded per image), or
                       // spi == 0: 32 consecutive pixels of the flat (image, pixel) axis -- the im2col walk wraps into the
                       // next image exactly as the channels-last top_diff's rows do, so nothing is padded
              int img, pix;
              if (p.spi > 0) { img = ks / p.spi; pix = (ks - img * p.spi) * BK; }
              else { const int hw = p.Ho * p.Wo, m0 = ks * BK; img = m0 / hw; pix = m0 - img * hw; }
              const int oh = pix / p.Wo;
              const int h = oh * p.sv - p.ph, w = (pix - oh * p.Wo) * p.sh - p.pw;
#pragma unroll
              for (int j = 0; j < kChunks; ++j) MNV_LDI(a_dst + j * 4096, &tmap_a, bar, a_c[j], w, h, img, a_kw[j], a_kh[j]);
            }
          }
          const uint32_t dst = smem_base + stage * kSBytes + kATile;
          const int halves = WIDE ? 2 : 1, rows = WIDE ? p.bn / 2 : p.bn;   // a TMA box has at most 256 rows
          if (dbg_skip & 2) {
          } else
          if (AM == A_TMA && p.b_im2col) {   // transposed orientation: the n-tile's bn pixels x 32 channels of tap (kh, kw)
            const int n_first = t.nt * p.bn, img = n_first / p.P, pix = n_first - img * p.P, oh = pix / p.Wo;
            const int tap = ks / p.cpt, cc = ks - tap * p.cpt, kh = tap / p.fw;
            tma_load_im2col_4d(dst, &tmap_b, full0 + 8 * stage, cc * BK, (pix - oh * p.Wo) * p.sh - p.pw, oh * p.sv - p.ph, img, tap - kh * p.fw, kh);   // never paired
          } else
          if (AM == A_TMA && p.b_mn) {   // MN-major B: one 32 n x 32 k box per 32 columns (bn % 32 == 0, never wide)
            if (p.b_mn == 2) MNV_LD3(dst, &tmap_b, fbar, 0, ks * BK, (t.nt * p.bn + n_rank) >> 5);
            else
            for (int j = 0; j < bn_cta / 32; ++j) MNV_LD2(dst + j * 4096, &tmap_b, fbar, t.nt * p.bn + n_rank + 32 * j, ks * BK);
          } else
          for (int h = 0; h < halves; ++h) {
            const uint32_t d2 = dst + h * rows * 128;
            const int n0 = t.nt * p.bn + n_rank + h * rows;
            if (p.spi > 0) {   // top_diff as (pixel, channel, image): one image's 32-pixel slab per stage
              int img = ks / p.spi;
              MNV_LD3(d2, &tmap_b, fbar, (ks - img * p.spi) * BK, n0, img);
            } else {
              MNV_LD2(d2, &tmap_b, fbar, ks * BK, n0);
            }
          }
          if (++stage == kNStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
#undef MNV_LD2
#undef MNV_LD3
#undef MNV_LDI
    // ===================== gather producers =====================
    if constexpr (AM == A_TMA) {
      // both operands come through TMA: no gather to do, help drain the accumulators
      epilogue(warp & 3, 1 + (warp - kProducerWarp0) / 4);
    } else if (BTMA) {
      // Stage-interleaved warp groups: group g (4 warps = 128 threads, one thread per tile row, all 32 k
      // of the stage) owns ring slot g, i.e. every 4th k-stage.  A group's loads for its next stage are
      // issued right after it hands the current one over and have three other stages' time to land, so
      // L2 latency is covered by thread-level parallelism rather than by one warp's scoreboard (a warp's
      // outstanding LDGs share scoreboard slots: register look-ahead inside one warp does not overlap).
      const int pw = warp - kProducerWarp0;                 // 0..15
      const int grp = pw >> 2;                              // ring slot owned by this group (groups >= kNStages idle)
      constexpr int ngrp = kNStages;
      const int gt = threadIdx.x - kProducerWarp0 * 32 - grp * 128;   // 0..127: tile row (MC) / chunk id (KC)
      const int kc_kq = gt & 7, kc_row0 = gt >> 3;          // KC mapping: rows kc_row0 + 16*i, chunk column kc_kq
      const uint32_t slot_full = full0 + 8 * grp, slot_empty = empty0 + 8 * grp;
      const uint32_t a_tile = smem_base + grp * kSBytes;
      uint32_t uses = 0;                                    // completed uses of this slot -> wait parity
      uint32_t cnt = 0;                                     // global k-stage counter at tile start
      float va[32];
      if constexpr (AM == A_IM2COL_FWD_K || AM == A_IM2COL_WGRAD_M) {
        // Lanes along the taps.  Warp wq of the group owns tile rows 32*wq .. 32*wq+31.
        //   FWD_K  : lane = k of the stage (its (offset, kh, kw) comes from the k-table, one LDS per stage); the warp
        //            walks its 32 rows, whose (base, kh-mask, kw-mask) live one per lane and are broadcast by SHFL;
        //            each lane stores one float per row (conflict-free: a row is 32 consecutive words).
        //   WGRAD_M: lane = tile row (tap), with a fixed source offset; the warp walks the 32 pixels of the stage,
        //            whose (base, masks) are computed one per lane per stage and broadcast; each lane then owns a whole
        //            128-byte row of the tile: eight conflict-free STS.128.
        const int wq = pw & 3;
        const int HW = p.H * p.W;
        for (int tile = tile_first; grp < ngrp && tile < total_tiles; tile += tile_step) {
          TileCoord t = decode_tile(p, tile);
          const int nks = t.ks_end - t.ks_begin;
          int ks = t.ks_begin + (grp + ngrp - static_cast<int>(cnt % static_cast<uint32_t>(ngrp))) % ngrp;
          cnt += static_cast<uint32_t>(nks);
          if (ks >= t.ks_end) continue;
          int my_off = 0;
          uint32_t my_h = 0, my_w = 0;      // FWD_K: this lane's row masks; WGRAD_M: this lane's kh, kw (31 = row past M)
          if (AM == A_IM2COL_FWD_K) {
            ARow ar = a_row_setup<A_IM2COL_FWD>(p, t.mt * BM + wq * 32 + lane);
            my_off = static_cast<int>(ar.base);
            my_h = ar.valid ? ar.mh : 0u; my_w = ar.mw;
          } else {
            const int m = t.mt * BM + wq * 32 + lane, ff = p.fh * p.fw;
            my_h = 31u; my_w = 0u;
            if (m < p.M) {
              const int ci = m / ff, rs = m - ci * ff, rr = rs / p.fw;
              my_h = static_cast<uint32_t>(p.fh - 1 - rr); my_w = static_cast<uint32_t>(p.fw - 1 - (rs - rr * p.fw));
              my_off = ci * HW + static_cast<int>(my_h) * p.W + static_cast<int>(my_w);
            }
          }
          auto load = [&](int k) {
            if (AM == A_IM2COL_FWD_K) {
              const uint32_t e = ld_shared_u32(ktab0 + 4 * (k * BK + lane));
              const int koff = static_cast<int>(e & 0x3FFFFFu);
              const uint32_t kh = (e >> 22) & 31u, kw = e >> 27;
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                const int rb = __shfl_sync(0xffffffffu, my_off, r);
                const uint32_t rh = __shfl_sync(0xffffffffu, my_h, r), rw = __shfl_sync(0xffffffffu, my_w, r);
                va[r] = (((rh >> kh) & (rw >> kw)) & 1u) ? __ldg(p.a + (rb + koff)) : 0.f;
              }
            } else {
              // this lane describes pixel `lane` of the stage: 32 pixels of image k / spi
              const int img = k / p.spi, pix = (k - img * p.spi) * BK + lane;
              int pb = 0;
              uint32_t ph_m = 0, pw_m = 0;
              if (pix < p.Ho * p.Wo) {
                const int oh = pix / p.Wo, h0 = oh * p.sv - p.ph, w0 = (pix - oh * p.Wo) * p.sh - p.pw;
                pb = img * p.Ci * HW + h0 * p.W + w0;
                for (int kh = 0; kh < p.fh; ++kh) if (h0 + kh >= 0 && h0 + kh < p.H) ph_m |= 1u << kh;
                for (int kw = 0; kw < p.fw; ++kw) if (w0 + kw >= 0 && w0 + kw < p.W) pw_m |= 1u << kw;
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int qb = __shfl_sync(0xffffffffu, pb, j);
                const uint32_t qh = __shfl_sync(0xffffffffu, ph_m, j), qw = __shfl_sync(0xffffffffu, pw_m, j);
                va[j] = (((qh >> my_h) & (qw >> my_w)) & 1u) ? __ldg(p.a + (qb + my_off)) : 0.f;
              }
            }
          };
          load(ks);
          for (; ks < t.ks_end; ks += ngrp) {
            mbar_wait(slot_empty, (uses & 1u) ^ 1u, p.wait_hint);
            if (AM == A_IM2COL_FWD_K) {
              const uint32_t lane_off = static_cast<uint32_t>(lane & 3) * 4u;
#pragma unroll
              for (int r = 0; r < 32; ++r)   // (32*wq + r) & 7 == r & 7
                st_shared_f32(a_tile + static_cast<uint32_t>(wq * 32 + r) * 128u + (static_cast<uint32_t>((lane >> 2) ^ (r & 7)) << 4) + lane_off,
                              to_tf32(va[r]));
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                st_shared_v4(a_tile + sw128_off(wq * 32 + lane, q), to_tf32(va[4 * q]), to_tf32(va[4 * q + 1]), to_tf32(va[4 * q + 2]),
                             to_tf32(va[4 * q + 3]));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(slot_full);
            ++uses;
            if (ks + ngrp < t.ks_end) load(ks + ngrp);
          }
        }
      } else
      for (int tile = tile_first; grp < ngrp && tile < total_tiles; tile += tile_step) {
        TileCoord t = decode_tile(p, tile);
        const int nks = t.ks_end - t.ks_begin;
        int ks = t.ks_begin + (grp + ngrp - static_cast<int>(cnt % static_cast<uint32_t>(ngrp))) % ngrp;   // this group's first stage in the tile
        cnt += static_cast<uint32_t>(nks);
        if (ks >= t.ks_end) continue;
        ARow arow = {};
        AWgrad<8> awg;
        if (AM == A_IM2COL_WGRAD) awg = a_wgrad_setup<8>(p, t.mt * BM, kc_row0, 16);
        else arow = a_row_setup<AM>(p, t.mt * BM + gt);
        const float* srcr = p.a + arow.base;
        uint32_t m0 = 0, m1 = 0;
        if (AM == A_IM2COL_FWD && p.use_ktab && arow.valid) {
          if (p.use_ktab == 1) {   // combined mask: bit kh*fw + kw
            for (int kh = 0; kh < p.fh; ++kh)
              if ((arow.mh >> kh) & 1u) m0 |= (arow.mw & ((1u << p.fw) - 1u)) << (kh * p.fw);
          } else {
            m0 = arow.mh; m1 = arow.mw;
          }
        }
        auto load = [&](int k) {
          if (AM == A_IM2COL_WGRAD) a_gather_wgrad<8>(p, awg, k, kc_kq, va);
          else if (AM == A_IM2COL_FWD && p.use_ktab == 1) a_gatherN_ktab<1, 32>(srcr, m0, m1, ktab0 + 4 * (k * BK), va);
          else if (AM == A_IM2COL_FWD && p.use_ktab == 2) a_gatherN_ktab<2, 32>(srcr, m0, m1, ktab0 + 4 * (k * BK), va);
          else a_gatherN<AM, 32>(p, arow, k * BK, va);
        };
        load(ks);
        for (; ks < t.ks_end; ks += ngrp) {
          mbar_wait(slot_empty, (uses & 1u) ^ 1u, p.wait_hint);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint32_t off = (AM == A_IM2COL_WGRAD) ? sw128_off(kc_row0 + 16 * q, kc_kq) : sw128_off(gt, q);
            st_shared_v4(a_tile + off, to_tf32(va[4 * q]), to_tf32(va[4 * q + 1]), to_tf32(va[4 * q + 2]), to_tf32(va[4 * q + 3]));
          }
          fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async proxy
          __syncwarp();
          if (lane == 0) mbar_arrive_relaxed(slot_full);
          ++uses;
          if (ks + ngrp < t.ks_end) load(ks + ngrp);   // lands while the other groups' stages are consumed
        }
      }
    } else {
    // register look-ahead in k-stages: the gathers are latency-bound on L2 when the ring is not full, so the
    // TMA-fed variants (8 staging registers per stage) look 4 stages ahead; gathered-B variants stage A and B.
    constexpr int LOOK = (AM == A_IM2COL_WGRAD) ? 1 : 2;
    const int pt = threadIdx.x - kProducerWarp0 * 32;   // 0..511
    const int a_row = pt & 127, a_q = pt >> 7;          // MC mapping: 8 consecutive k: [a_q*8, +8)
    const int b_kq = pt & 7, b_row0 = pt >> 3;          // KC mapping: rows b_row0 + 64*i, chunk column b_kq
    const int b_iters = (p.bn + 63) / 64;               // <= 4
    int stage = 0; uint32_t phase = 0;
    float va[LOOK][8];
    float4 vb[BTMA ? 1 : LOOK][4];
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      TileCoord t = decode_tile(p, tile);
      ARow arow = {};
      AWgrad<2> awg;
      if (AM == A_IM2COL_WGRAD) awg = a_wgrad_setup<2>(p, t.mt * BM, b_row0, 64);
      else arow = a_row_setup<AM>(p, t.mt * BM + a_row);
      // k-table path: byte pointer of the row's window origin + validity mask(s)
      const float* srcr = p.a + arow.base;
      uint32_t m0 = 0, m1 = 0;
      if (AM == A_IM2COL_FWD && p.use_ktab && arow.valid) {
        if (p.use_ktab == 1) {   // combined mask: bit kh*fw + kw
          for (int kh = 0; kh < p.fh; ++kh)
            if ((arow.mh >> kh) & 1u) m0 |= (arow.mw & ((1u << p.fw) - 1u)) << (kh * p.fw);
        } else {
          m0 = arow.mh; m1 = arow.mw;
        }
      }
      const int n_base = t.nt * p.bn;
      auto load = [&](int l, int ks) {
        if (AM == A_IM2COL_WGRAD) a_gather_wgrad<2>(p, awg, ks, b_kq, va[l]);
        else if (AM == A_IM2COL_FWD && p.use_ktab == 1) a_gatherN_ktab<1, 8>(srcr, m0, m1, ktab0 + 4 * (ks * BK + a_q * 8), va[l]);
        else if (AM == A_IM2COL_FWD && p.use_ktab == 2) a_gatherN_ktab<2, 8>(srcr, m0, m1, ktab0 + 4 * (ks * BK + a_q * 8), va[l]);
        else a_gatherN<AM, 8>(p, arow, ks * BK + a_q * 8, va[l]);
        if (!BTMA) b_gather<BMD>(p, n_base, b_row0, b_iters, ks * BK + b_kq * 4, vb[BTMA ? 0 : l]);
      };
      auto store = [&](int l) {
        const uint32_t a_tile = smem_base + stage * kSBytes;
        const uint32_t b_tile = a_tile + kATile;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t off = (AM == A_IM2COL_WGRAD) ? sw128_off(b_row0 + 64 * q, b_kq) : sw128_off(a_row, a_q * 2 + q);
          st_shared_v4(a_tile + off, to_tf32(va[l][4 * q]), to_tf32(va[l][4 * q + 1]), to_tf32(va[l][4 * q + 2]), to_tf32(va[l][4 * q + 3]));
        }
        if (!BTMA) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i >= b_iters) break;
            int row = b_row0 + 64 * i;
            if (row < p.bn) {
              const float4& s4 = vb[BTMA ? 0 : l][i];
              st_shared_v4(b_tile + sw128_off(row, b_kq), to_tf32(s4.x), to_tf32(s4.y), to_tf32(s4.z), to_tf32(s4.w));
            }
          }
        }
      };
#pragma unroll
      for (int l = 0; l < LOOK; ++l)
        if (t.ks_begin + l < t.ks_end) load(l, t.ks_begin + l);
      for (int ks = t.ks_begin; ks < t.ks_end; ks += LOOK) {
#pragma unroll
        for (int l = 0; l < LOOK; ++l) {
          if (ks + l < t.ks_end) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1, p.wait_hint);
            store(l);
            fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(full0 + 8 * stage);
            if (++stage == kNStages) { stage = 0; phase ^= 1; }
            if (ks + l + LOOK < t.ks_end) load(l, ks + l + LOOK);  // flies while the ring drains
          }
        }
      }
    }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();     // neither CTA frees TMEM the pair's last MMAs may still write
  if (warp == 4) { if (PAIR) tmem_dealloc_2sm(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols); }
}

#undef tile_first
#undef tile_step
#undef bn_cta

// split-K: out[index(m,n)] = bias[n] + sum_s partial[s][n][m], splits folded in order
__global__ void __launch_bounds__(kBlock) splitk_reduce_kernel(const GemmParams p) {
  pdl_enter();
  size_t mn = static_cast<size_t>(p.M) * p.N;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < mn;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int m = static_cast<int>(t % p.M), n = static_cast<int>(t / p.M);
    if (!out_row_ok(p, m)) continue;
    float acc = __ldg(p.partial + t);
    for (int s = 1; s < p.splits; ++s) acc += __ldg(p.partial + static_cast<size_t>(s) * mn + t);
    if (p.bias) acc += __ldg(p.bias + n);
    if (p.relu) acc = acc > 0.f ? acc : 0.f;
    p.out[out_index(p, m, n)] = acc;
  }
}

// tail split: the same fold for the K-split tiles of the last wave only (tile ids >= tail_first in raster order)
__global__ void __launch_bounds__(kBlock) splitk_tail_reduce_kernel(const GemmParams p) {
  pdl_enter();
  const int tile_m = p.tall ? 2 * BM : BM, per_tile = tile_m * p.bn;
  const int ntail = p.m_tiles * p.n_tiles - p.tail_first;
  const size_t mn = static_cast<size_t>(p.M) * p.N, total = static_cast<size_t>(ntail) * per_tile;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int tt = static_cast<int>(i / per_tile), e = static_cast<int>(i - static_cast<size_t>(tt) * per_tile);
    const TileCoord t = decode_tile(p, p.tail_first + tt);   // split 0 of that tile: same (mt, nt)
    const int m = t.mt * tile_m + e % tile_m, n = t.nt * p.bn + e / tile_m;
    if (n >= p.N || !out_row_ok(p, m)) continue;
    const size_t at = static_cast<size_t>(n) * p.M + m;
    float acc = __ldg(p.partial + at);
    for (int sp = 1; sp < p.tail_splits; ++sp) acc += __ldg(p.partial + static_cast<size_t>(sp) * mn + at);
    if (p.bias) acc += __ldg(p.bias + n);
    if (p.relu) acc = acc > 0.f ? acc : 0.f;
    p.out[out_index(p, m, n)] = acc;
  }
}

// w[co][ci][r][s] -> wt[ci][co][r'][s']: backward-data reads the filter as B[n=ci][k=(co,r,s)].
// flip = 1 additionally rotates the taps by 180 degrees (r' = fh-1-r, s' = fw-1-s), which turns a
// stride-1 backward-data into a forward convolution of top_diff with pad' = f-1-pad.
__global__ void __launch_bounds__(kBlock) filter_swap_kernel(const float* __restrict__ w, float* __restrict__ wt, int Co, int Ci, int ff, int flip, int round) {
  pdl_enter();
  size_t total = static_cast<size_t>(Co) * Ci * ff;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int rs = static_cast<int>(t % ff);
    size_t rest = t / ff;
    int co = static_cast<int>(rest % Co), ci = static_cast<int>(rest / Co);
    // rounded to TF32 here because the tensor core only truncates what TMA delivers
    float v = __ldg(w + (static_cast<size_t>(co) * Ci + ci) * ff + (flip ? ff - 1 - rs : rs));
    wt[t] = round ? to_tf32(v) : v;
  }
}

// ------------------------------------------------------------------------------------------------
// shift-GEMM convolution: stride-1 cross-correlation straight out of a shared-memory copy of the input
// ------------------------------------------------------------------------------------------------
// A channels-last activation xv[pos][C] (pos = (n*Hv + Y)*Wv + X, flat over the whole batch) is its own im2col
// matrix up to a ROW SHIFT: for the output at flat position pos, tap (a, b) reads row pos + a*Wv + b.  So one
// "plane" of 256 + halo consecutive rows x 32 channels (128 B rows, 128B swizzle, one TMA load) serves every tap of
// a 256-position tile: the tap only moves the UMMA descriptor's start address by whole 128 B rows.  The 128B swizzle
// is a function of the absolute shared-memory address on both the TMA and the UMMA side, so a start address that is
// not 1024 B aligned needs nothing else (measured with tools/shift_diag.py: the base-offset field must stay 0).
// Nothing is duplicated on the L2 -> SM path, which is what bounds the im2col-fed kernel on narrow outputs
// (AlexNet conv1 after space-to-depth: 3.0 GB of TMA traffic as im2col, 0.3 GB as planes).
// Positions whose window would leave the image (X >= Wo or Y >= Ho) are computed and dropped by the epilogue
// (conv1: 57x57 positions per 55x55 outputs = 7 % more tensor work).
// Orientation: D[filter][position] -- the filters are the UMMA's M (128 rows, Co <= 128 of them real), the 256
// positions its N.  One 128 x 256 x 8 instruction costs ~120 cycles whatever N is filled with; with the positions as
// M, a 96-filter output needs 128 x 96 x 8 instructions of ~95 cycles each (measured: 50 % pipe-active but 2.2x the
// time).  The epilogue transposes 32 x 32 blocks through shared memory so that a warp stores 32 consecutive pixels of
// one filter plane (128 B).  k-stage order is chunk-major (all taps of channel chunk 0, then chunk 1, ...) so a plane
// is released as soon as its taps are issued and the 3-plane ring prefetches across tiles.
struct ShiftParams {
  const float* bias;     // per filter, may be null
  float* out;            // out[n][co][Ho*Wo]
  int Co, bn, Ho, Wo, Hv, Wv, fwv, taps, cpt, relu;
  int nk_last;           // 8-channel k-steps of the last channel chunk (1..4): all-zero steps are not issued
  int b_stages;          // depth of the filter ring (<= kShBStagesMax)
  int m_tiles;           // ceil(total positions / 256)
  int plane_boxes;       // 128-row TMA boxes per plane: ceil((256 + halo) / 128) <= 3
  int total_pos;         // N * Hv * Wv
  int dbg;               // tuning experiments: bit 0 = no global stores, bit 1 = no epilogue body at all
  unsigned wait_hint;
};
constexpr int kShTileM = 256;
constexpr int kShPlaneRows = 384, kShPlaneBytes = kShPlaneRows * 128, kShPlanes = 3;
constexpr int kShBStagesMax = 8;
constexpr int kShEpiWarps = 8, kShMmaWarp = 8, kShPlaneWarp = 9 /* warp 10: filter producer */, kShThreads = 352;
constexpr int kShStagingBytes = kShEpiWarps * 32 * 32 * 4;   // one 32 x 32 transpose block per epilogue warp
constexpr int kShSmemMax = 232448;   // 227 KB opt-in limit
constexpr int kShSmemFixed = kShPlanes * kShPlaneBytes + kShStagingBytes + 1024 + 256;

// k-steps (8 channels of one tap) run chunk-major, then tap, then 8-channel group, skipping the groups of the last chunk
// that lie wholly past the real channels (nk_last of 4 remain).  A filter stage (bn x 32 floats) holds one tap of a full
// chunk, or 4 / nk_last taps of the last one, so its 128 B rows may mix taps (AlexNet conv1 view: 48 channels =
// 9 stages of 4 steps + 5 stages of 2 x 2 steps, not 18 stages).
__global__ void __launch_bounds__(kShThreads, 1) conv_shift_fwd_kernel(const __grid_constant__ ShiftParams p,
                                                                        const __grid_constant__ CUtensorMap tmap_x,
                                                                        const __grid_constant__ CUtensorMap tmap_w) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t b_bytes = static_cast<uint32_t>(p.bn) * 128u;
  const uint32_t plane0 = smem_u32(smem), bring0 = plane0 + kShPlanes * kShPlaneBytes;
  float* staging = reinterpret_cast<float*>(smem + kShPlanes * kShPlaneBytes + p.b_stages * b_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kShPlanes * kShPlaneBytes + p.b_stages * b_bytes + kShStagingBytes);
  const uint32_t pfull0 = smem_u32(bars), pempty0 = smem_u32(bars + 3), bfull0 = smem_u32(bars + 6), bempty0 = smem_u32(bars + 6 + kShBStagesMax);
  const uint32_t tfull0 = smem_u32(bars + 6 + 2 * kShBStagesMax), tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10 + 2 * kShBStagesMax);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tps_last = 4 / p.nk_last;   // taps per filter stage in the last channel chunk
  const int nstages = (p.cpt - 1) * p.taps + (p.taps + tps_last - 1) / tps_last;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kShPlanes; ++i) { mbar_init(pfull0 + 8 * i, 1); mbar_init(pempty0 + 8 * i, 1); }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(bfull0 + 8 * i, 1); mbar_init(bempty0 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, kShEpiWarps); }
    fence_barrier_init();
  }
  if (warp == kShMmaWarp) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kShEpiWarps) {
    // ===================== epilogue: TMEM (lane = filter, column = position) -> 32 x 32 transpose in shared memory
    // -> (+bias, relu) -> out, lane = position =====================
    const int quarter = warp & 3, chalf = warp >> 2;
    const int P = p.Ho * p.Wo, HWv = p.Hv * p.Wv;
    const int nco = min(32, p.Co - quarter * 32);        // <= 0: this warp's TMEM lanes hold no real filter
    const uint32_t blk = smem_u32(staging) + warp * 4096;   // explicit shared-space accesses: a generic pointer here compiles to LD.E / ST.E
    // bias and ReLU are applied on the TMEM side, where a thread is one filter: one scalar per thread for the whole kernel
    const float bias = (p.bias && lane < nco) ? __ldg(p.bias + quarter * 32 + lane) : 0.f;
    const bool relu = p.relu != 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      mbar_wait(tfull0 + 8 * acc, acc_phase, p.wait_hint);
      tc_fence_after();
      if (nco > 0 && !(p.dbg & 2)) {
        for (int cb = 0; cb < 4; ++cb) {
          const int col0 = chalf * 128 + cb * 32;
          uint32_t r0[16], r1[16];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * 256 + col0);
          tmem_ld16(taddr, r0);
          tmem_ld16(taddr + 16, r1);
#pragma unroll
          for (int j = 0; j < 16; ++j) {   // element (filter = lane, position = j) at lane * 32 + (j ^ lane): conflict-free both ways
            const float v0 = __uint_as_float(r0[j]) + bias, v1 = __uint_as_float(r1[j]) + bias;
            st_shared_f32(blk + 4 * (lane * 32 + (j ^ lane)), (relu && !(v0 > 0.f)) ? 0.f : v0);
            st_shared_f32(blk + 4 * (lane * 32 + ((j + 16) ^ lane)), (relu && !(v1 > 0.f)) ? 0.f : v1);
          }
          __syncwarp();
          const int pos = tile * kShTileM + col0 + lane;
          const int n = pos / HWv, rem = pos - n * HWv, Y = rem / p.Wv, X = rem - Y * p.Wv;
          const bool ok = pos < p.total_pos && Y < p.Ho && X < p.Wo;
          if (ok && !(p.dbg & 1)) {
            float* dst = p.out + (static_cast<size_t>(n) * p.Co + quarter * 32) * P + Y * p.Wo + X;
#pragma unroll 8
            for (int l = 0; l < nco; ++l) dst[static_cast<size_t>(l) * P] = __uint_as_float(ld_shared_u32(blk + 4 * (l * 32 + (lane ^ l))));
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp == kShMmaWarp) {
    // ===================== MMA issuer: one thread; descriptor high words are constants, low words = address >> 4 ======
    if (lane == 0) {
      const uint32_t idesc = make_idesc(kShTileM);
      const uint32_t desc_hi = static_cast<uint32_t>(make_sw128_desc(0) >> 32);
      const uint32_t w_lo0 = (bring0 & 0x3FFFFu) >> 4, w_step = b_bytes >> 4;
      const uint32_t p_lo0 = (plane0 & 0x3FFFFu) >> 4, row_step = static_cast<uint32_t>(p.Wv) * 8u;
      const int cpt = p.cpt, taps = p.taps, fwv = p.fwv, nk_last = p.nk_last, nbs = p.b_stages;
      int ps = 0, bs = 0, acc = 0; uint32_t pph = 0, bph = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1, p.wait_hint);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * 256;
        uint32_t accum = 0;
        for (int chunk = 0; chunk < cpt; ++chunk) {
          const int nk = chunk + 1 == cpt ? nk_last : 4, tps = 4 / nk;   // taps per filter stage
          mbar_wait(pfull0 + 8 * ps, pph, p.wait_hint);
          tc_fence_after();
          uint32_t row_lo = p_lo0 + ps * (kShPlaneBytes >> 4), x_lo = row_lo;
          int b = 0;
          for (int tap = 0; tap < taps;) {
            mbar_wait(bfull0 + 8 * bs, bph, p.wait_hint);
            tc_fence_after();
            uint32_t w_lo = w_lo0 + bs * w_step;
            for (int g = 0; g < tps && tap < taps; ++g, ++tap) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (kk < nk) {   // A = filters (rows past bn read the next stage's bytes: those accumulator lanes are never drained)
                  umma_tf32_lohi(d0, w_lo, x_lo + kk * 2, desc_hi, idesc, accum);
                  accum = 1;
                  w_lo += 2;
                }
              }
              x_lo += 8u;
              if (++b == fwv) { b = 0; row_lo += row_step; x_lo = row_lo; }
            }
            umma_commit(bempty0 + 8 * bs);
            if (++bs == nbs) { bs = 0; bph ^= 1; }
          }
          umma_commit(pempty0 + 8 * ps);
          if (++ps == kShPlanes) { ps = 0; pph ^= 1; }
        }
        umma_commit(tfull0 + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == kShPlaneWarp) {
    // ===================== plane producer: (256 + halo) rows x 32 channels per (tile, chunk) =====================
    if (lane == 0) {
      int ps = 0; uint32_t pph = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        for (int chunk = 0; chunk < p.cpt; ++chunk) {
          mbar_wait(pempty0 + 8 * ps, pph ^ 1, p.wait_hint);
          mbar_arrive_expect_tx(pfull0 + 8 * ps, static_cast<uint32_t>(p.plane_boxes) * 128u * 128u);
          for (int j = 0; j < p.plane_boxes; ++j)   // rows past the tensor arrive as zeros
            tma_load_2d(plane0 + ps * kShPlaneBytes + j * 128 * 128, &tmap_x, pfull0 + 8 * ps, chunk * BK, tile * kShTileM + j * 128);
          if (++ps == kShPlanes) { ps = 0; pph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== filter producer: one bn x 32 stage per four k-steps =====================
    if (lane == 0) {
      int bs = 0; uint32_t bph = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        for (int ks = 0; ks < nstages; ++ks) {
          mbar_wait(bempty0 + 8 * bs, bph ^ 1, p.wait_hint);
          mbar_arrive_expect_tx(bfull0 + 8 * bs, b_bytes);
          tma_load_2d(bring0 + bs * b_bytes, &tmap_w, bfull0 + 8 * bs, ks * BK, 0);
          if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kShMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Tuning / debug options.  The PRODUCT library (libmnv_b200.so) is built without MNV_TUNING: every option is a
// compile-time constant, there is no process-global mutable state and neither mnv_debug_set_option nor the
// SIMT checker kernel exists in it.  The tuning build (libmnv_b200_tuning.so, -DMNV_TUNING; include/mnv_debug.h) is what
// tools/ and the alternate-path parity tests load.
#ifdef MNV_TUNING
struct Opt {
  std::atomic<int> v;
  constexpr Opt(int d) : v(d) {}
  int load() const { return v.load(std::memory_order_relaxed); }
  int exchange(int x) { return v.exchange(x); }
};
#define MNV_OPT static Opt
#else
struct Opt {
  int d;
  constexpr Opt(int d_) : d(d_) {}
  constexpr int load() const { return d; }
};
#define MNV_OPT static constexpr Opt
#endif
MNV_OPT g_opt_pf_dist{0};      // TMA L2-prefetch distance in k-stages (0 = off)
MNV_OPT g_opt_wait_hint{100};  // mbarrier.try_wait suspend hint in ns (tuning)
MNV_OPT g_opt_no_wide{0};     // 1: never use the wide (bn > 256) tile (tuning)
MNV_OPT g_opt_simt{0};       // 1: run the SIMT checker instead of tcgen05 (debug only)
MNV_OPT g_opt_max_splits{0}; // >0: clamp split-K (debug / tuning)
MNV_OPT g_opt_no_tma{0};     // 1: gather B with threads even where TMA applies (debug)
MNV_OPT g_opt_no_fwd_bwd{0};  // 1: use the generic backward-data gather for stride 1 too (debug)
MNV_OPT g_opt_no_tma_a{0};   // bit 0: no TMA-im2col fprop/dgrad, bit 1: no TMA MatMult A, bit 2: no TMA wgrad (debug / tuning)
MNV_OPT g_opt_tma_tf32{1};   // operand maps typed TFLOAT32: TMA then rounds fp32 -> tf32 to nearest on the way in (measured:
                                             // norm-rel error vs fp64 2.9e-4 unbiased, against 7.7e-4 with a -7e-4 bias for FLOAT32 maps,
                                             // whose low mantissa bits the tensor core just drops); 0 = FLOAT32 maps (debug)
MNV_OPT g_opt_force_tma_a{0}; // 1: take the all-TMA conv path whenever it applies, ignoring the profitability rule (tuning)
MNV_OPT g_opt_no_klane{0};   // 1: strided convs keep the lanes-along-pixels gathers (debug / tuning)
MNV_OPT g_opt_no_deep{0};    // 1: keep the 4 x 48 KB ring for bn <= 128 on the all-TMA path (tuning)
MNV_OPT g_opt_no_tall{0};    // 1: never use the 256-row tile (tuning)
MNV_OPT g_opt_tall_min_stages{32};  // shortest per-tile mainloop (k-stages) the 256-row tile is used for (measured on AlexNet conv3-5:
                                    // 32 beats 64 by 3-7 % on forward / backward-data, 16 and 8 are no better; FC GEMMs do not care)
MNV_OPT g_opt_no_ktab{0};    // 1: table-free forward gather (debug)
MNV_OPT g_opt_no_s2d{0};     // 1: strided few-channel convs stay on the gather kernel (debug / tuning)
MNV_OPT g_opt_shift_dbg{0};  // shift-GEMM kernel experiments (see ShiftParams::dbg)
MNV_OPT g_opt_no_shift{0};   // 1: no shift-GEMM kernel (debug / tuning)
MNV_OPT g_opt_no_transposed{0};  // 1: narrow-output convolutions keep the D[pixel][co] orientation (tuning)
// The CTA pair (RING 4, tcgen05.mma.cta_group::2) is built, bit-identical to the other tiles and OFF: measured on B200 it is
// slower than the 256-row single-CTA tile it would replace -- MatMult 8192^3 845 TF/s (tall) vs 744 (pair, the peer's copies
// signalling the leader's barrier, the CUTLASS protocol) vs 686 (local barriers + a forwarded arrive); conv4 forward 0.233 ms vs
// 0.314; conv4 backward-filter 0.315 vs 0.534.  With tf32 an instruction covers K = 8 (32 bytes), so a k-stage is four paired
// instructions plus per-copy / per-stage cross-SM signalling; the halved B bytes do not buy that back.  DESIGN.md 5.1d.
MNV_OPT g_opt_wgrad_wide{1};    // 0: channels-last backward-filter never takes the wide (128 x 384) tile (tuning)
MNV_OPT g_opt_no_pointwise{0};  // 1: 1x1 convolutions through the im2col map and the packed filter like every other geometry (tuning)
MNV_OPT g_opt_pair_remote{1};   // CTA pair: 1 = the peer's copies count their bytes on the leader's barrier (cta_group::2 copy forms),
                                // 0 = every copy signals its own CTA's barrier and the peer forwards one arrive per stage
MNV_OPT g_opt_pair{0};          // 1: 256-row tiles run on a CTA pair (RING 4) instead of one CTA (RING 3) (tuning)
MNV_OPT g_opt_no_mn3{0};        // 1: MN-major operands as one 2-D box per 32 rows instead of one 3-D box per k-stage (tuning)
MNV_OPT g_opt_no_nhwc_wgrad{0}; // 1: backward-filter keeps the re-pitched NCHW top_diff (K-major B) instead of the channels-last one (tuning)
MNV_OPT g_opt_s2d_im2col{0}; // 1: space-to-depth views may also run on the im2col-fed kernel (experiments; slower than the gathers)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static std::atomic<EncodeTiledFn> cached{nullptr};
  EncodeTiledFn fn = cached.load(std::memory_order_acquire);
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  cached.store(fn, std::memory_order_release);
  return fn;
}
// Operands staged by a pre-pass are rounded to TF32 there only when the tensor maps are plain FLOAT32 (rounding
// twice would double-round); the SIMT checker reads them unrounded.
static int prepass_round() { return (g_opt_tma_tf32.load() || g_opt_simt.load()) ? 0 : 1; }
static CUtensorMapDataType tmap_dtype() {
  return g_opt_tma_tf32.load() ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
}
// K-major fp32 matrix [rows][ld] seen as a 2-D tensor (k fastest); box = 32 k x bn rows, 128B swizzle.
static bool make_b_tmap(CUtensorMap* tm, const float* b, int rows, int K, int ld, int bn) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * sizeof(float)};
  cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(bn)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, tmap_dtype(), 2, const_cast<float*>(b), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// column-major fp32 matrix [K][ld] (m fastest) as a 2-D tensor; box = 32 m x 32 k, 128B swizzle: the MN-major
// operand slab of one 32-row chunk.
static bool make_a_mn_tmap(CUtensorMap* tm, const float* a, int M, int K, int ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(M), static_cast<cuuint64_t>(K)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * sizeof(float)};
  cuuint32_t box[2] = {32, BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, tmap_dtype(), 2, const_cast<float*>(a), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// The same matrix as a (32 m, K, m / 32) 3-D tensor: one box of `groups` slabs = the `groups` consecutive 32 m x 32 k MN-major
// slabs of a k-stage, laid out in shared memory exactly as `groups` boxes of make_a_mn_tmap at 4 KB steps -- one TMA
// instruction instead of `groups` (the single producer thread's issue rate bounds the MN-major kernels).  A slab's 32
// columns are read without a bound on m, so the caller guarantees round32(M) <= ld (rows past M are never stored).
static bool make_mn3_tmap(CUtensorMap* tm, const float* a, int M, int K, int ld, int groups) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn || groups < 1 || groups > 12 || (M + 31) / 32 * 32 > ld) return false;
  cuuint64_t dims[3] = {32, static_cast<cuuint64_t>(K), static_cast<cuuint64_t>((M + 31) / 32)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * sizeof(float), 32 * sizeof(float)};
  cuuint32_t box[3] = {32, BK, static_cast<cuuint32_t>(groups)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, tmap_dtype(), 3, const_cast<float*>(a), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeIm2colFn get_im2col_fn() {
  static std::atomic<EncodeIm2colFn> cached{nullptr};
  EncodeIm2colFn fn = cached.load(std::memory_order_acquire);
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) return nullptr;
  fn = reinterpret_cast<EncodeIm2colFn>(ptr);
  cached.store(fn, std::memory_order_release);
  return fn;
}
// Channels-last activation copy x[N][H][W][Cp] as a (C, W, H, N) im2col tensor.  The bounding box of the filter
// window's origin is [-pad, dim + pad - (f-1)) per spatial axis, traversed with the convolution stride; one load
// is 32 channels x `pixels` window positions (rows of 128 B, 128B swizzle).
static bool make_im2col_tmap(CUtensorMap* tm, const float* x, int Cp, int W, int H, int N, int pw, int ph, int fw, int fh, int sh,
                             int sv, int pixels, bool mn_major) {
  EncodeIm2colFn fn = get_im2col_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(Cp), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(Cp) * 4, static_cast<cuuint64_t>(W) * Cp * 4, static_cast<cuuint64_t>(H) * W * Cp * 4};
  int lower[2] = {-pw, -ph};
  int upper[2] = {pw - (fw - 1), ph - (fh - 1)};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(sh), static_cast<cuuint32_t>(sv), 1};
  CUresult r = fn(tm, tmap_dtype(), 4, const_cast<float*>(x), dims, strides, lower, upper, BK, static_cast<cuuint32_t>(pixels), estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  // Drivers up to CUDA 13.1 mis-encode im2col maps of tensors smaller than 128 KB (same correction as CUTLASS's
  // make_im2col_tma_copy_desc applies).
  int drv = 0;
  if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 &&
      static_cast<unsigned long long>(N) * H * W * Cp * 4ull < 131072ull)
    reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
  return true;
}

// x[n][c][hw] -> y[n][hw][Cp]: channels-last copy for the im2col tensor maps, channels C..Cp-1 zero (rounding to
// TF32 is done by the TFLOAT32-typed tensor map).  Tiles go through shared memory so both sides stay coalesced.
// RELU: the source is ReLU backward's result computed on the fly, v = act > 0 ? x : 0, also written in NCHW to `dx` (the
// reference op's output): one pass (8 B read + 8 B written per element) instead of ReLU backward (12 B) plus this copy (8 B).
// Pixels per tile = 32 * TWJ, chosen per plane size (nhwc_tile_j) so that the last tile of a plane is not mostly padding:
// 13 x 13 = 169 pixels are one tile of 192 (88 % useful) instead of 128 + 41 (66 %), 14 x 14 one of 224, 28 x 28 five of 160.
static int nhwc_tile_j(int HW) {
  int best = 4;
  long long best_cover = -1;
  for (int j = 2; j <= 8; ++j) {
    const long long tw = 32 * j, cover = (HW + tw - 1) / tw * tw;
    if (best_cover < 0 || cover < best_cover || (cover == best_cover && j > best)) { best = j; best_cover = cover; }
  }
  return best;
}
template <bool RELU, int TWJ>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int Cp, int HW,
                                                           int tiles_c, int tiles_hw, long long total_tiles, float* __restrict__ tilesum,
                                                           const float* __restrict__ act, float* __restrict__ dx) {
  pdl_enter();
  // tile = 32 channels x 32 * TWJ pixels: 4 * TWJ independent 128-byte-per-warp loads per thread before the barrier.
  // tilesum != null: the pass also leaves the sum of every (image, pixel tile, channel) in tilesum[(n * tiles_hw + th) * C + c]
  // -- for a top_diff that is ConvBackwardBias's per-tile partial, folded over (n, th) by rowsum_fold_kernel in a fixed order.
  constexpr int TW = 32 * TWJ;
  __shared__ float tile[32][TW + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int tc = static_cast<int>(t % tiles_c);
    const long long r = t / tiles_c;
    const int th = static_cast<int>(r % tiles_hw);
    const size_t n = static_cast<size_t>(r / tiles_hw);
    const int c0 = tc * 32, h0 = th * TW;
    const float* src = x + (n * C + c0) * HW + h0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = wid + 8 * i;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < TWJ; ++j) {
        const int hw = lane + 32 * j;
        float v = 0.f;
        if (c0 + c < C && h0 + hw < HW) {
          const size_t o = static_cast<size_t>(c) * HW + hw;
          v = __ldg(src + o);
          if (RELU) {
            const size_t g = (n * C + c0) * HW + h0 + o;
            v = __ldg(act + g) > 0.f ? v : 0.f;
            dx[g] = v;
          }
        }
        tile[c][hw] = v;
        acc += v;
      }
      if (tilesum) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0 && c0 + c < C) tilesum[(n * tiles_hw + th) * C + c0 + c] = acc;
      }
    }
    __syncthreads();
    if (c0 + lane < Cp) {
      float* dst = y + (n * HW + h0) * Cp + c0 + lane;
#pragma unroll
      for (int i = 0; i < 4 * TWJ; ++i) {
        const int hw = wid + 8 * i;
        if (h0 + hw < HW) dst[static_cast<size_t>(hw) * Cp] = tile[lane][hw];
      }
    }
    __syncthreads();
  }
}
template <bool RELU>
static int launch_nhwc_t(const float* x, float* y, int N, int C, int Cp, int HW, cudaStream_t s, float* tilesum, const float* act, float* dx) {
  const int twj = nhwc_tile_j(HW), tiles_c = (Cp + 31) / 32, tiles_hw = (HW + 32 * twj - 1) / (32 * twj);
  const long long total = static_cast<long long>(N) * tiles_c * tiles_hw, cap = static_cast<long long>(kNumSMs) * 16;
  const dim3 grid(static_cast<unsigned>(total < cap ? total : cap));
#define MNV_NHWC(J) case J: launch_pdl((nchw_to_nhwc_kernel<RELU, J>), grid, dim3(256), 0, s, x, y, C, Cp, HW, tiles_c, tiles_hw, total, tilesum, act, dx); break;
  switch (twj) { MNV_NHWC(2) MNV_NHWC(3) MNV_NHWC(4) MNV_NHWC(5) MNV_NHWC(6) MNV_NHWC(7) default: MNV_NHWC(8) }
#undef MNV_NHWC
  return finish_launch();
}
// B operand of the TMA-im2col convolutions: out[n][tap][c] (c < Kc = 32-channel chunks per tap, zero past C)
//   = w[n*sn + c*sc + (flip ? ff-1-tap : tap)], TF32-rounded.
__global__ void __launch_bounds__(kBlock) filter_pack_kernel(const float* __restrict__ w, float* __restrict__ out, int rows, int C, int ff,
                                                             int Kc, long long sn, long long sc, int flip, int round) {
  pdl_enter();
  size_t total = static_cast<size_t>(rows) * ff * Kc;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int c = static_cast<int>(t % Kc);
    size_t rest = t / Kc;
    int tap = static_cast<int>(rest % ff);
    size_t n = rest / ff;
    float v = c < C ? __ldg(w + n * sn + static_cast<size_t>(c) * sc + (flip ? ff - 1 - tap : tap)) : 0.f;
    out[t] = round ? to_tf32(v) : v;
  }
}

// ---- space-to-depth view of strided few-channel convolutions (AlexNet conv1: 3 channels, 11x11 / 4) -----------------
// A stride-(sv,sh) cross-correlation over Ci channels is a stride-1 one over Ci*sv*sh channels:
//   xv[n][Y][X][(c,dy,dx)] = x[n][c][Y*sv + dy - ph][X*sh + dx - pw]          (zero outside the image)
//   wv[co][(a,b)][(c,dy,dx)] = w[co][c][a*sv + dy][b*sh + dx]                 (zero past the real filter)
//   y[oy][ox] = sum_{a,b,(c,dy,dx)} xv[oy + a][ox + b][.] * wv[(a,b)][.],    a < ceil(fh/sv), b < ceil(fw/sh)
// xv is channels-last with >= 32 channels, so the im2col tensor maps feed the tensor core (no gather warps, whose
// 12-k-stage tiles were latency-bound: 128 TF/s) at the price of multiplying some zeros (AlexNet conv1: K 363 -> 576).
struct S2D {
  int Ci, H, W, ph, pw, sv, sh, fh, fw;   // the real convolution
  int Civ, Hv, Wv, fhv, fwv;              // the stride-1 view (pad 0)
};
static bool s2d_plan(S2D* v, int Ci, int H, int W, int Ho, int Wo, int ph, int pw, int sv, int sh, int fh, int fw) {
  if (sv == 1 && sh == 1) return false;
  v->Ci = Ci; v->H = H; v->W = W; v->ph = ph; v->pw = pw; v->sv = sv; v->sh = sh; v->fh = fh; v->fw = fw;
  v->Civ = Ci * sv * sh; v->fhv = (fh + sv - 1) / sv; v->fwv = (fw + sh - 1) / sh;
  v->Hv = Ho + v->fhv - 1; v->Wv = Wo + v->fwv - 1;
  if (Ci >= BK || v->Civ > 1024) return false;                            // enough channels already: the direct map is better
  const int cpt = (v->Civ + BK - 1) / BK;
  if (cpt * BK * 2 > v->Civ * 3) return false;                            // same padding rule as conv_tma_fprop
  if (static_cast<size_t>(Ci) * sv * (static_cast<size_t>(v->Wv) * sh + 8) * sizeof(float) > 48 * 1024) return false;   // staging rows
  return true;
}
// One CTA per (image, group of R virtual rows): the R * Ci*sv source rows are staged in shared memory (coalesced reads
// along x, all loads of the group in flight before the barrier), then the R * Wv * Cp output floats are written in
// order (coalesced), channels Civ..Cp-1 zero.  VEC4 (sh == 4, Civ % 4 == 0): one 16-byte store = 4 consecutive x of
// one staged row.
__host__ __device__ inline int s2d_pitch(const S2D& v) { return (v.Wv * v.sh + 3) / 4 * 4 + 4; }
template <bool VEC4>
__global__ void __launch_bounds__(256) s2d_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, const S2D v, int Cp, int R, int groups) {
  pdl_enter();
  extern __shared__ float4 rows4[];
  float* rows = reinterpret_cast<float*>(rows4);
  const int L = v.Wv * v.sh, pitch = s2d_pitch(v), SR = v.Ci * v.sv;
  const int g = blockIdx.x % groups, Y0 = g * R, Rc = min(R, v.Hv - Y0);
  const size_t n = blockIdx.x / groups;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int rr = wid; rr < Rc * SR; rr += 8) {
    const int yl = rr / SR, r = rr - yl * SR, c = r / v.sv, yy = (Y0 + yl) * v.sv + (r - c * v.sv) - v.ph;
    const bool row_ok = yy >= 0 && yy < v.H;
    const float* src = x + ((n * v.Ci + c) * v.H + (row_ok ? yy : 0)) * v.W;
    // 4-byte cp.async with zero fill: nothing waits on a register, so every load of the group is in flight at once
    // (an LDG loop whose trip count is not a compile-time constant runs its tail one load at a time)
    const uint32_t dst_row = smem_u32(rows + rr * pitch);
    for (int i = lane; i < L; i += 32) {
      const int xx = i - v.pw;
      const bool ok = row_ok && xx >= 0 && xx < v.W;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_row + 4 * i), "l"(src + (ok ? xx : 0)), "r"(ok ? 4 : 0) : "memory");
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  float* dst = y + (n * v.Hv + Y0) * static_cast<size_t>(v.Wv) * Cp;
  if (VEC4) {   // Cq consecutive threads write one pixel's Cp floats; no division inside the loops
    const int Cq = Cp >> 2, ppi = 256 / Cq, q = threadIdx.x % Cq, p0 = threadIdx.x / Cq;
    if (p0 < ppi) {
      const bool live = 4 * q < v.Civ;
      float4* dst4 = reinterpret_cast<float4*>(dst);
      for (int yl = 0; yl < Rc; ++yl) {
        const float* src = rows + (yl * SR + q) * pitch;   // q = staged row (sh == 4)
#pragma unroll 4
        for (int X = p0; X < v.Wv; X += ppi)
          dst4[(yl * v.Wv + X) * Cq + q] = live ? *reinterpret_cast<const float4*>(src + X * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  } else {
    const int per_row = v.Wv * Cp, total = Rc * per_row;
    for (int o = threadIdx.x; o < total; o += 256) {
      const int yl = o / per_row, rem = o - yl * per_row, X = rem / Cp, ch = rem - X * Cp, r = ch / v.sh;
      dst[o] = ch < v.Civ ? rows[(yl * SR + r) * pitch + X * v.sh + (ch - r * v.sh)] : 0.f;
    }
  }
}
// B operand of the view, out[n][k] with row length Kp floats, zero past Civ and past the real filter:
//   shift == 0 (im2col-fed kernel): k = (tap * cpt + chunk) * 32 + j, channel ch = chunk * 32 + j;
//   shift != 0 (shift-GEMM kernel): k = stage * 32 + slot * 8 + j in the kernel's stage order -- a stage is one tap of a
//              full channel chunk (slot = 8-channel group kk), or 4 / nk_last taps of the last chunk (slot = t * nk_last + kk);
//              ch = chunk * 32 + kk * 8 + j.
//   value = w[n*sn + c*sc + (flip ? ff-1-t : t)], t = (a*sv + dy) * fw + b*sh + dx, ch = (c*sv + dy)*sh + dx.
__global__ void __launch_bounds__(kBlock) s2d_filter_pack_kernel(const float* __restrict__ w, float* __restrict__ out, int rows, const S2D v,
                                                                 int cpt, int shift, int nk_last, int Kp, long long sn, long long sc, int flip,
                                                                 int round) {
  pdl_enter();
  const int ffv = v.fhv * v.fwv, ff = v.fh * v.fw;
  size_t total = static_cast<size_t>(rows) * Kp;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(t % Kp);
    const size_t n = t / Kp;
    int tap, ch;
    bool live = true;
    if (!shift) {
      const int ks = k / BK, chunk = ks % cpt;
      tap = ks / cpt; ch = chunk * BK + k % BK;
    } else {
      const int stage = k / BK, slot = (k % BK) >> 3, full = (cpt - 1) * ffv;
      if (stage < full) {
        tap = stage % ffv; ch = (stage / ffv) * BK + slot * 8 + (k & 7);
      } else {
        const int tps = 4 / nk_last, tin = slot / nk_last;
        tap = (stage - full) * tps + tin; ch = (cpt - 1) * BK + (slot - tin * nk_last) * 8 + (k & 7);
        live = tin < tps && tap < ffv;
      }
    }
    float val = 0.f;
    if (live) {
      const int a = tap / v.fwv, b = tap - a * v.fwv, r = ch / v.sh, c = r / v.sv;
      const int kh = a * v.sv + (r - c * v.sv), kw = b * v.sh + (ch - r * v.sh);
      if (ch < v.Civ && kh < v.fh && kw < v.fw) {
        const int tr = kh * v.fw + kw;
        val = __ldg(w + n * sn + static_cast<size_t>(c) * sc + (flip ? ff - 1 - tr : tr));
      }
    }
    out[t] = round ? to_tf32(val) : val;
  }
}
static int launch_s2d(const float* x, float* y, int N, const S2D& v, int Cp, cudaStream_t s) {
  const size_t row_bytes = static_cast<size_t>(v.Ci) * v.sv * s2d_pitch(v) * sizeof(float);
  int R = static_cast<int>(48 * 1024 / row_bytes);
  R = R > 4 ? 4 : R < 1 ? 1 : R;
  const int groups = (v.Hv + R - 1) / R;
  const unsigned grid = static_cast<unsigned>(N) * groups;
  if (v.sh == 4 && v.Civ % 4 == 0 && aligned16(y)) launch_pdl((s2d_nhwc_kernel<true>), dim3(grid), dim3(256), R * row_bytes, s, x, y, v, Cp, R, groups);
  else launch_pdl((s2d_nhwc_kernel<false>), dim3(grid), dim3(256), R * row_bytes, s, x, y, v, Cp, R, groups);
  return finish_launch();
}

// top_diff[img][co][pitch] (pitch % 4 == 0, only the first P pixels of a row are real) as a 3-D tensor;
// box = 32 pixels x bn channels x 1 image.  Pixels past P are out of bounds => zero-filled.
static bool make_dy_tmap(CUtensorMap* tm, const float* dy, int P, int pitch, int Co, int N, int bn) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(P), static_cast<cuuint64_t>(Co), static_cast<cuuint64_t>(N)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(pitch) * sizeof(float), static_cast<cuuint64_t>(pitch) * Co * sizeof(float)};
  cuuint32_t box[3] = {BK, static_cast<cuuint32_t>(bn), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, tmap_dtype(), 3, const_cast<float*>(dy), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// row-pitch repack for TMA: dst[r][0..inner) = src[r][0..inner), dst rows `pitch` floats apart.  rowsum != null: the
// pass also leaves each row's sum (of the unrounded values) in rowsum[r] -- for top_diff rows (image, channel) that is
// the per-image bias gradient, folded over the images by rowsum_fold_kernel (ConvBackwardBias for free: the 4 B per
// element its own kernel would re-read are already in registers here).
__global__ void __launch_bounds__(kBlock) repitch_kernel(const float* __restrict__ src, float* __restrict__ dst, int inner, int pitch, size_t rows, int round,
                                                         float* __restrict__ rowsum) {
  pdl_enter();
  // one warp per row keeps both sides coalesced without integer division
  size_t warp = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  size_t nwarps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
  int lane = threadIdx.x & 31;
  for (size_t r = warp; r < rows; r += nwarps) {
    const float* s = src + r * inner;
    float* d = dst + r * pitch;
    float acc = 0.f;
    if (round) for (int i = lane; i < inner; i += 32) { const float v = __ldg(s + i); d[i] = to_tf32(v); acc += v; }
    else for (int i = lane; i < inner; i += 32) { const float v = __ldg(s + i); d[i] = v; acc += v; }
    if (rowsum) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) rowsum[r] = acc;
    }
  }
}
// db[c] = sum over the rows of rowsum[n * C + c] (rows = images, or (image, pixel tile) pairs).  A CTA folds 32 channels:
// 32 threads per channel take every 32nd row (independent loads, 128 B per warp), then the 32 partial sums are added in lane
// order -- a fixed order, so the result is deterministic.
__global__ void __launch_bounds__(1024) rowsum_fold_kernel(const float* __restrict__ rowsum, float* __restrict__ db, int N, int C) {
  pdl_enter();
  // 32 channels x 32 row lanes per CTA; every lane keeps 4 independent loads in flight (the first version, 8 row lanes with
  // one CTA per 32 channels, was latency-bound: 16 us per call for 1536 rows)
  __shared__ float part[32][33];
  const int cl = threadIdx.x & 31, nl = threadIdx.x >> 5, c = blockIdx.x * 32 + cl;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < C) {
    int n = nl;
    for (; n + 96 < N; n += 128) {
      a0 += __ldg(rowsum + static_cast<size_t>(n) * C + c);
      a1 += __ldg(rowsum + static_cast<size_t>(n + 32) * C + c);
      a2 += __ldg(rowsum + static_cast<size_t>(n + 64) * C + c);
      a3 += __ldg(rowsum + static_cast<size_t>(n + 96) * C + c);
    }
    for (; n < N; n += 32) a0 += __ldg(rowsum + static_cast<size_t>(n) * C + c);
  }
  part[nl][cl] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (nl == 0 && c < C) {
    float acc = part[0][cl];
#pragma unroll
    for (int i = 1; i < 32; ++i) acc += part[i][cl];
    db[c] = acc;
  }
}

// SMs the persistent tensor-core kernel may occupy.  It runs one CTA per SM for the whole launch, so when another
// persistent kernel (an overlapped NCCL all-reduce) holds a few SMs, a 148-CTA grid runs its last CTAs as a second
// wave and the launch takes up to twice as long; the data-parallel trainer therefore leaves NCCL's SMs out
// ("sm_budget" option; default = all 148).
MNV_OPT g_opt_sm_budget{kNumSMs};
static int sm_budget() {
  int v = g_opt_sm_budget.load();
  return v < 1 ? 1 : v > kNumSMs ? kNumSMs : v;
}
// split-K for a tile grid that does not fill the machine (needs the caller's workspace for the partials)
static void plan_splits(GemmParams& p, size_t ws_bytes_for_partials) {
  long long tiles = static_cast<long long>(p.m_tiles) * p.n_tiles;
  int splits = 1;
  const int sms = p.tall == 2 ? sm_budget() / 2 : sm_budget();   // CTA pairs: one tile per two SMs
  if (tiles * 4 < sms * 3) {  // fewer than 3/4 of a wave: split K
    splits = static_cast<int>(sms / tiles);
    int max_by_k = p.k_stages / 4;  // at least 4 stages per split
    if (splits > max_by_k) splits = max_by_k;
    size_t per_split = static_cast<size_t>(p.M) * p.N * sizeof(float);
    size_t max_by_ws = per_split ? ws_bytes_for_partials / per_split : 0;
    if (static_cast<size_t>(splits) > max_by_ws) splits = static_cast<int>(max_by_ws);
    int clamp = g_opt_max_splits.load();
    if (clamp > 0 && splits > clamp) splits = clamp;
    if (splits < 1) splits = 1;
  }
  p.stages_per_split = (p.k_stages + splits - 1) / splits;
  p.splits = (p.k_stages + p.stages_per_split - 1) / p.stages_per_split;  // no empty split
}
// fraction of the machine's CTA slots the tile grid keeps busy over its whole run
MNV_OPT g_opt_no_tail{0};    // 1: no tail split (tuning)
// tail_ok: the caller K-splits a thin last wave over the idle CTAs (plan_tail), so that wave costs 1 / S of a wave plus
// the partial round trip (taken as a tenth of a wave)
static double wave_fill(const GemmParams& p, bool tail_ok = false) {
  long long ctas = static_cast<long long>(p.m_tiles) * p.n_tiles * p.splits;
  const int sms = p.tall == 2 ? sm_budget() / 2 : sm_budget();
  long long waves = (ctas + sms - 1) / sms;
  const int rem = static_cast<int>(ctas % sms);
  if (tail_ok && p.splits == 1 && ctas >= sms && rem > 0 && rem * 2 <= sms && !g_opt_no_tail.load()) {
    int S = sms / rem;
    if (S > 8) S = 8;
    if (S > p.k_stages / 8) S = p.k_stages / 8;
    if (S >= 2) return (static_cast<double>(ctas) / sms) / (static_cast<double>(ctas / sms) + 1.0 / S + 0.1);
  }
  return static_cast<double>(ctas) / static_cast<double>(waves * sms);
}

static void plan_tiles(GemmParams& p, size_t ws_bytes_for_partials, bool allow_wide = false, bool allow_tall = false, bool tail_ok = false) {
  p.m_tiles = (p.M + BM - 1) / BM;
  p.wide = 0; p.tall = 0; p.stages = kStages; p.stage_bytes = kStageBytes;
  int n_tiles, bn;
  if (allow_wide && p.N > BN_MAX && !g_opt_no_wide.load()) {
    // 256 < N: tiles of up to 384 columns (two UMMA halves share one gathered A tile) in a 3-deep ring of
    // 64 KB slots -- the same 192 KB of shared memory as the 4 x 48 KB ring of the narrow tile
    constexpr int kWideMax = 384;
    n_tiles = (p.N + kWideMax - 1) / kWideMax;
    bn = (p.N + n_tiles - 1) / n_tiles;
    bn = (bn + 31) / 32 * 32;
    if (bn > BN_MAX) { p.wide = 1; p.stages = 3; p.stage_bytes = kABytes + kWideMax * BK * 4; }
  } else {
    n_tiles = (p.N + BN_MAX - 1) / BN_MAX;
    bn = (p.N + n_tiles - 1) / n_tiles;
    bn = (bn + 15) / 16 * 16;
  }
  if (bn < 16) bn = 16;
  p.bn = bn;
  p.n_tiles = (p.N + bn - 1) / bn;
  p.k_stages = (p.K + BK - 1) / BK;
  if (p.k_stages < 1) p.k_stages = 1;
  plan_splits(p, ws_bytes_for_partials);
  // The wide tile has a single TMEM accumulator, so its epilogue is not hidden behind the next tile's
  // mainloop: worth it only when a tile's mainloop is long (measured break-even ~70 k-stages).
  if (p.wide && p.stages_per_split < 96) plan_tiles(p, ws_bytes_for_partials, false, allow_tall, tail_ok);
  if (allow_tall && !g_opt_no_tall.load()) {
    // 256-row tile (all-TMA path): halves the B traffic per flop.  Like the wide tile it has no second accumulator
    // set, so it wants a long mainloop, and it must not cost machine fill.
    GemmParams q = p;
    q.tall = 1; q.wide = 0; q.stages = 3; q.stage_bytes = 2 * kABytes + kBBytes;
    q.m_tiles = (p.M + 2 * BM - 1) / (2 * BM);
    q.n_tiles = (p.N + BN_MAX - 1) / BN_MAX;
    q.bn = ((p.N + q.n_tiles - 1) / q.n_tiles + 15) / 16 * 16;
    if (q.bn < 16) q.bn = 16;
    q.n_tiles = (p.N + q.bn - 1) / q.bn;
    plan_splits(q, ws_bytes_for_partials);
    if (q.stages_per_split >= g_opt_tall_min_stages.load() && (wave_fill(q, tail_ok) >= 0.85 || wave_fill(q, tail_ok) >= wave_fill(p, tail_ok))) {
      p = q;
      // the same 256-row tile on a CTA pair (cta_group::2): each CTA stages half of B, so bn is a multiple of 32
      if (g_opt_pair.load() && sm_budget() % 2 == 0) {
        GemmParams r = q;
        r.bn = (q.bn + 31) / 32 * 32;
        if (r.bn <= BN_MAX) {
          r.n_tiles = (r.N + r.bn - 1) / r.bn;
          r.tall = 2; r.stages = 6; r.stage_bytes = kABytes + 128 * BK * 4;
          plan_splits(r, ws_bytes_for_partials);
          p = r;
        }
      }
    }
  }
}

// Tail split (see GemmParams): when the tile grid ends in a last wave that fills at most half the machine, the tiles of
// that wave are K-split over the idle CTAs.  Partials use the split-K layout partial[split][n][m] (only the tail tiles'
// entries are touched); splitk_tail_reduce_kernel folds them.
static void plan_tail(GemmParams& p, void* ws, size_t ws_bytes) {
  p.tail_splits = 0;
  if (g_opt_no_tail.load() || p.splits != 1 || !ws) return;
  const int sms = p.tall == 2 ? sm_budget() / 2 : sm_budget(), tiles = p.m_tiles * p.n_tiles;
  const int rem = tiles % sms;
  if (tiles < sms || rem == 0 || rem * 2 > sms) return;
  int S = sms / rem;
  if (S > 8) S = 8;
  if (S > p.k_stages / 8) S = p.k_stages / 8;     // at least 8 k-stages per tail split
  if (S < 2) return;
  if (ws_bytes < static_cast<size_t>(S) * p.M * p.N * sizeof(float)) return;
  p.tail_stages = (p.k_stages + S - 1) / S;
  p.tail_splits = (p.k_stages + p.tail_stages - 1) / p.tail_stages;
  p.tail_first = tiles - rem;
  p.partial = static_cast<float*>(ws);
}
static const CUtensorMap& null_tmap() { static CUtensorMap z = {}; return z; }
template <int AM, int BMD, bool BTMA, int RING>
static int launch_umma_w(const GemmParams& p, const CUtensorMap& tm, cudaStream_t s, const CUtensorMap& tm_a = null_tmap()) {
  // opt in to >48 KB dynamic shared memory once per (device, instantiation)
  static std::atomic<uint64_t> attr_done{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_done.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(umma_gemm_kernel<AM, BMD, BTMA, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  GemmParams q = p;
  const int tiles = p.m_tiles * p.n_tiles;
  if (q.tail_splits > 1 && q.splits == 1 && q.partial) {
    q.total_units = q.tail_first + (tiles - q.tail_first) * q.tail_splits;
  } else {
    q.tail_splits = 0; q.tail_first = 0x7fffffff; q.total_units = tiles * p.splits;
  }
  const long long total = q.total_units;
  const int sms = sm_budget();
  if (RING == 4) {   // CTA pairs: one cluster of 2 per tile in flight
    const int pairs = static_cast<int>(total < sms / 2 ? total : sms / 2);
    launch_pdl_cluster((umma_gemm_kernel<AM, BMD, BTMA, RING>), dim3(2 * pairs), dim3(kThreads), kSmemBytes, s, 2, q, tm, tm_a);
    return finish_launch();
  }
  int grid = static_cast<int>(total < sms ? total : sms);
  launch_pdl((umma_gemm_kernel<AM, BMD, BTMA, RING>), dim3(grid), dim3(kThreads), kSmemBytes, s, q, tm, tm_a);
  return finish_launch();
}
template <int AM, int BMD, bool BTMA>
static int launch_umma(const GemmParams& p, const CUtensorMap& tm, cudaStream_t s) {
  if (BTMA && p.wide) return launch_umma_w<AM, BMD, BTMA, BTMA>(p, tm, s);   // the wide tile exists on the TMA-fed path only
  return launch_umma_w<AM, BMD, BTMA, false>(p, tm, s);
}

// both operands through TMA
static int launch_umma_tma(const GemmParams& p, const CUtensorMap& tm_a, const CUtensorMap& tm_b, cudaStream_t s) {
  int rc = p.tall == 2 ? launch_umma_w<A_TMA, B_KMAJOR, true, 4>(p, tm_b, s, tm_a)
           : p.tall    ? launch_umma_w<A_TMA, B_KMAJOR, true, 3>(p, tm_b, s, tm_a)
           : p.wide    ? launch_umma_w<A_TMA, B_KMAJOR, true, 1>(p, tm_b, s, tm_a)
           : (p.bn <= 128 && !g_opt_no_deep.load()) ? launch_umma_w<A_TMA, B_KMAJOR, true, 2>(p, tm_b, s, tm_a)
                          : launch_umma_w<A_TMA, B_KMAJOR, true, 0>(p, tm_b, s, tm_a);
  if (rc) return rc;
  if (p.splits == 1 && p.tail_splits > 1 && p.partial) {
    const size_t items = static_cast<size_t>(p.m_tiles * p.n_tiles - p.tail_first) * (p.tall ? 2 * BM : BM) * p.bn;
    launch_pdl(splitk_tail_reduce_kernel, dim3(stream_grid(items)), dim3(kBlock), 0, s, p);
    return finish_launch();
  }
  if (p.splits == 1) return rc;
  launch_pdl(splitk_reduce_kernel, dim3(stream_grid(static_cast<size_t>(p.M) * p.N)), dim3(kBlock), 0, s, p);
  return finish_launch();
}

template <int AM, int BMD>
static int launch_gemm(GemmParams& p, void* ws, size_t ws_bytes, cudaStream_t s) {
#ifdef MNV_TUNING
  if (g_opt_simt.load()) {
    p.splits = 1; p.partial = nullptr;
    simt_gemm_kernel<AM, BMD><<<stream_grid(static_cast<size_t>(p.M) * p.N), 256, 0, s>>>(p);
    return finish_launch();
  }
#endif
  // A K-major B whose rows are not 16-byte pitched/aligned (conv1: K = 363; odd GEMM k) is re-pitched into the
  // workspace once (B is the small operand: filters, or the batch-side matrix of an FC layer) so that it can
  // still come through TMA and the kernel can run the grouped-producer path.
  if (BMD == B_KMAJOR && !p.b_vec && ws && !g_opt_no_tma.load()) {
    int pitch = (p.K + 3) / 4 * 4;
    size_t need = (static_cast<size_t>(p.N) * pitch * sizeof(float) + 255) / 256 * 256;
    if (ws_bytes >= need) {
      float* packed = static_cast<float*>(ws);
      launch_pdl(repitch_kernel, dim3(stream_grid(static_cast<size_t>(p.N) * 32)), dim3(kBlock), 0, s, p.b, packed, p.K, pitch, static_cast<size_t>(p.N), prepass_round(), nullptr);
      int rc0 = finish_launch();
      if (rc0) return rc0;
      p.b = packed; p.ldb = pitch; p.b_vec = 1;
      ws = static_cast<uint8_t*>(ws) + need;
      ws_bytes -= need;
    }
  }
  const bool tma_ok = BMD == B_KMAJOR && p.b_vec && !g_opt_no_tma.load() && get_encode_fn() != nullptr;
  // column-major A is an MN-major UMMA operand as it lies in memory: with B on TMA too there are no gather warps
  const bool tma_a_ok = AM == A_COLMAJOR && tma_ok && !(g_opt_no_tma_a.load() & 2) && p.lda % 4 == 0 && aligned16(p.a);
  plan_tiles(p, ws ? ws_bytes : 0, tma_ok, tma_a_ok);
  p.use_ktab = 0;
  if (AM == A_IM2COL_FWD && p.k_stages * BK <= kKtabMax && !g_opt_no_ktab.load()) {
    long long img = static_cast<long long>(p.Ci) * p.H * p.W;
    if (p.fh * p.fw <= 31 && img < (1ll << 27)) p.use_ktab = 1;
    else if (img < (1ll << 22)) p.use_ktab = 2;
  }
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  bool tma = false;
  if (tma_ok) {
    tma = make_b_tmap(&tm, p.b, p.N, p.K, p.ldb, (p.wide || p.tall == 2) ? p.bn / 2 : p.bn);
    if (!tma && (p.wide || p.tall)) plan_tiles(p, ws ? ws_bytes : 0, false);   // the wide / tall tiles need the TMA-fed path
  }
  p.partial = p.splits > 1 ? static_cast<float*>(ws) : nullptr;
  int rc;
  if (tma_a_ok && tma) {
    CUtensorMap tm_a;
    memset(&tm_a, 0, sizeof(tm_a));
    if (!g_opt_no_mn3.load() && make_mn3_tmap(&tm_a, p.a, p.M, p.K, p.lda, p.tall == 1 ? 8 : 4)) {
      p.a_mode = TMA_A_TILED_MN; p.a_g3 = 1;
      return launch_umma_tma(p, tm_a, tm, s);
    }
    if (make_a_mn_tmap(&tm_a, p.a, p.M, p.K, p.lda)) {
      p.a_mode = TMA_A_TILED_MN;
      return launch_umma_tma(p, tm_a, tm, s);
    }
  }
  if (p.tall) {   // planned for the all-TMA path, which did not materialise
    plan_tiles(p, ws ? ws_bytes : 0, tma);
    p.partial = p.splits > 1 ? static_cast<float*>(ws) : nullptr;
    if (tma && !make_b_tmap(&tm, p.b, p.N, p.K, p.ldb, (p.wide || p.tall == 2) ? p.bn / 2 : p.bn)) return MNV_EINVAL;
  }
  if (AM == A_IM2COL_FWD && BMD == B_KMAJOR && tma && (p.sv > 1 || p.sh > 1) && !g_opt_no_klane.load() &&
      p.k_stages * BK <= kKtabMax && static_cast<long long>(p.Ci) * p.H * p.W < (1ll << 22) && !p.wide) {
    p.use_ktab = 2;   // strided forward conv: lanes along k (taps are contiguous along kw, output pixels are not)
    rc = launch_umma<A_IM2COL_FWD_K, B_KMAJOR, true>(p, tm, s);
  } else if (BMD == B_KMAJOR && tma) rc = launch_umma<AM, B_KMAJOR, true>(p, tm, s);
  else rc = launch_umma<AM, BMD, false>(p, tm, s);
  if (rc || p.splits == 1) return rc;
  launch_pdl(splitk_reduce_kernel, dim3(stream_grid(static_cast<size_t>(p.M) * p.N)), dim3(kBlock), 0, s, p);
  return finish_launch();
}

static void zero_conv(GemmParams& p) {
  p.Ci = p.Co = p.H = p.W = p.Ho = p.Wo = p.fh = p.fw = 1;
  p.ph = p.pw = 0; p.sv = p.sh = 1;
  p.lda = p.ldb = 0; p.b_vec = 0; p.use_ktab = 0; p.spi = 0; p.a_mode = 0; p.b_mn = 0; p.cpt = 1; p.out_mode = 0; p.relu = 0; p.pf_dist = g_opt_pf_dist.load(); p.tail_first = 0; p.tail_splits = 0; p.tail_stages = 0; p.total_units = 0; p.r_ci = p.r_fh = p.r_fw = p.r_sv = p.r_sh = 1; p.b_flip = 0; p.b_im2col = 0; p.a_g3 = 0; p.pair_remote = g_opt_pair_remote.load(); p.wait_hint = static_cast<unsigned>(g_opt_wait_hint.load());
  p.bias = nullptr; p.partial = nullptr;
}

static int check_conv(int N, int Ci, int Co, int H, int W, int ph, int pw, int sv, int sh, int fh, int fw) {
  if (N < 0 || Ci <= 0 || Co <= 0 || H <= 0 || W <= 0 || ph < 0 || pw < 0 || sv <= 0 || sh <= 0 || fh <= 0 || fw <= 0)
    return MNV_EINVAL;
  if (fh > 31 || fw > 31) return MNV_EUNSUPPORTED;  // validity masks are 32-bit, bit 31 is the k-table sentinel
  if (H + 2 * ph < fh || W + 2 * pw < fw) return MNV_EINVAL;
  return MNV_OK;
}
static bool fits_int(long long v) { return v > 0 && v < 0x7fffffffLL; }

static size_t round256(size_t v) { return (v + 255) / 256 * 256; }
static int launch_nhwc(const float* x, float* y, int N, int C, int Cp, int HW, cudaStream_t s, float* tilesum = nullptr) {
  return launch_nhwc_t<false>(x, y, N, C, Cp, HW, s, tilesum, nullptr, nullptr);
}
static int nhwc_tiles_hw(int HW) { const int tw = 32 * nhwc_tile_j(HW); return (HW + tw - 1) / tw; }

// Channels-last "twin" of an NCHW activation, owned by the CALLER (mnv_conv_twin_bytes): several calls of one training
// step want the same copy (forward and backward-filter read x, backward-data and backward-filter read top_diff), so the
// *_tw entry points take (pointer, in/out state) and whichever call comes first fills it.  Without a twin the copy goes
// to the workspace, once per call, as before.
struct Twin {
  float* ptr;        // null: none
  int* state;        // *state != 0: already filled
  bool usable() const { return ptr != nullptr && state != nullptr; }
  bool valid() const { return usable() && (*state & 1) != 0; }
  bool has_sums() const { return valid() && (*state & 2) != 0; }   // the per-tile channel sums behind the copy are filled too
};
// Layout of a twin buffer of mnv_conv_twin_bytes(): [n][hw][Cp] floats, then (256-byte aligned) the per-(image, pixel tile, channel)
// sums a filling pass may leave for ConvBackwardBias: tilesum[(n * tiles_hw + th) * C + c], tiles of 32 * nhwc_tile_j(HW) pixels.
static size_t twin_copy_bytes(int N, int C, int HW) { return round256(static_cast<size_t>(N) * HW * ((C + 3) / 4 * 4) * sizeof(float)); }
static size_t twin_sums_bytes(int N, int C, int HW) { return round256(static_cast<size_t>(N) * nhwc_tiles_hw(HW) * C * sizeof(float)); }

// Cross-correlation of x[N][Ci][H][W] with `ff` taps as a GEMM whose two operands both arrive through TMA:
//   A[m = (n,oh,ow)][k = (tap, c)]  channels-last copy of x through an im2col tensor map (no gather warps, padding
//                                   is the map's out-of-bounds zero fill),
//   B[n = co][k = (tap, c)]         = w[co*w_sn + c*w_sc + (flip ? ff-1-tap : tap)], packed once per call.
// Both copies are TF32-rounded pre-passes into the workspace (one streaming pass each, << the GEMM).
// *done stays false when the path does not apply (tiny channel counts, no workspace): the caller falls back to
// the gather kernel.
static int conv_tma_fprop(const float* x, const float* w, long long w_sn, long long w_sc, int flip, const float* bias, int relu, float* out, int N,
                          int Ci, int Co, int H, int W, int Ho, int Wo, int ph, int pw, int sv, int sh, int fh, int fw, void* ws,
                          size_t ws_bytes, cudaStream_t s, bool* done, const S2D* s2d = nullptr, Twin twin = Twin{nullptr, nullptr}) {
  // s2d != null: x and w are the REAL strided convolution's operands (w strides w_sn / w_sc over its real taps) and the
  // geometry arguments describe its stride-1 space-to-depth view; only the two pre-passes differ.
  *done = false;
  if ((g_opt_no_tma_a.load() & 1) || g_opt_simt.load() || g_opt_no_tma.load() || !ws) return MNV_OK;
  const int cpt = (Ci + BK - 1) / BK, ff = fh * fw;
  if (cpt * BK * 2 > Ci * 3) return MNV_OK;                       // > 1.5x channel padding: the gather kernel wastes less
  if (ph > 127 || pw > 127 || fh - 1 - ph > 128 || fw - 1 - pw > 128 || sv > 8 || sh > 8) return MNV_OK;
  if (!get_im2col_fn() || !get_encode_fn()) return MNV_OK;
  const int Cp = (Ci + 3) / 4 * 4;
  const long long M = static_cast<long long>(N) * Ho * Wo, K = static_cast<long long>(ff) * cpt * BK;
  if (!fits_int(M) || !fits_int(K) || !fits_int(static_cast<long long>(N) * H * W * Cp)) return MNV_OK;
  const bool use_twin = twin.usable() && !s2d;
  const size_t x_bytes = use_twin ? 0 : round256(static_cast<size_t>(N) * H * W * Cp * sizeof(float));
  const size_t b_bytes = round256(static_cast<size_t>(Co) * K * sizeof(float));
  if (ws_bytes < x_bytes + b_bytes) return MNV_OK;
  float* xh = use_twin ? twin.ptr : static_cast<float*>(ws);
  float* wb = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + x_bytes);
  void* ws2 = static_cast<uint8_t*>(ws) + x_bytes + b_bytes;
  const size_t ws2_bytes = ws_bytes - x_bytes - b_bytes;

  GemmParams p;
  zero_conv(p);
  p.Ci = Ci; p.Co = Co; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.fh = fh; p.fw = fw; p.ph = ph; p.pw = pw; p.sv = sv; p.sh = sh;
  p.a = xh; p.b = wb; p.bias = bias; p.out = out;
  p.M = static_cast<int>(M); p.N = Co; p.K = static_cast<int>(K);
  p.ldb = p.K; p.b_vec = 1;
  p.P = Ho * Wo; p.img_stride = static_cast<long long>(Co) * p.P; p.col_stride = p.P;
  p.a_mode = TMA_A_IM2COL_K; p.cpt = cpt; p.relu = relu;
  // 1x1 / stride 1 / no padding (two thirds of GoogLeNet's convolutions): the channels-last copy IS the K-major A matrix
  // [pixel][Cp] -- a tiled map instead of the im2col one -- and the filter as stored is already a TMA-able B operand:
  // forward reads w[co][c] as K-major rows (Ci % 4 == 0), backward-data reads the same array as the MN-major B[k = co][n = ci]
  // (one 3-D box per k-stage, make_mn3_tmap; Ci % 32 == 0).  No filter pre-pass, no launch for it.
  const bool pointwise = !s2d && ff == 1 && sv == 1 && sh == 1 && ph == 0 && pw == 0 && !g_opt_no_pointwise.load();
  int b_direct = 0;
  if (pointwise && aligned16(w) && prepass_round() == 0) {     // (the TFLOAT32-typed maps round the operands as they copy)
    if (w_sc == 1 && w_sn == Ci && Ci % 4 == 0) b_direct = 1;
    else if (w_sn == 1 && w_sc == Co && Co % 32 == 0 && !g_opt_no_mn3.load()) b_direct = 2;
  }
  plan_tiles(p, ws2_bytes, b_direct != 2, true, true);
  if (b_direct == 2) {      // MN-major B: boxes of 32 columns per CTA
    p.bn = (p.bn + 31) / 32 * 32;
    if (p.bn > BN_MAX) p.bn = BN_MAX;
    p.n_tiles = (p.N + p.bn - 1) / p.bn;
    plan_splits(p, ws2_bytes);
  }
  // Round 1 took this path only when the GEMM did enough work per input element to pay for the channels-last pre-pass
  // (Co * taps >= 3000, or a narrow output).  With the twins shared between the calls of a step (and between the branches
  // of an inception module) the pre-pass is paid once per array, and the rule lost on both nets: always taken now
  // (GoogLeNet b120 15.95 -> 14.56 ms per step, AlexNet b256 5.66 -> 5.63 ms).  "no_tma_a" still forces the gather kernel.
  (void)g_opt_force_tma_a;
  p.partial = p.splits > 1 ? static_cast<float*>(ws2) : nullptr;
  plan_tail(p, ws2, ws2_bytes);
  CUtensorMap tm_a, tm_b;
  memset(&tm_a, 0, sizeof(tm_a));
  memset(&tm_b, 0, sizeof(tm_b));
  // Narrow outputs (80..128 filters: AlexNet conv2 backward-data, GoogLeNet's 96 / 112 / 128-filter layers): a 128 x 96 x 8
  // UMMA costs ~95 cycles against 138 for 128 x 256 x 8, so D[pixel][co] runs the tensor pipe at ~0.43 of its rate however
  // well it is fed.  Transposed, D[co][pixel], the filters are the 128-lane operand (TMA_A_TILED_K over the packed filter)
  // and 256 pixels the N of every instruction (the im2col box becomes the B tile): ~0.75 for 96 filters.  The epilogue then
  // owns 16 consecutive pixels of one output plane per thread.  Only when the pixel tiles fill the machine without split-K.
  bool transposed = false;
  GemmParams q = p;
  if (!s2d && Co >= 80 && Co <= BM && !g_opt_no_transposed.load()) {
    q.M = Co; q.N = static_cast<int>(M); q.a = wb; q.b = xh; q.lda = p.K; q.a_mode = TMA_A_TILED_K; q.b_im2col = 1; q.out_mode = 3;
    q.tail_first = 0; q.tail_splits = 0; q.tail_stages = 0; q.partial = nullptr;
    plan_tiles(q, 0, false, false);
    transposed = q.splits == 1 && q.m_tiles == 1 && q.n_tiles >= sm_budget() / 2 &&
                 make_b_tmap(&tm_a, wb, Co, p.K, p.K, BM) && make_im2col_tmap(&tm_b, xh, Cp, W, H, N, pw, ph, fw, fh, sh, sv, q.bn, false);
  }
  bool pack_filter = true;
  if (!transposed) {
    if (pointwise && make_b_tmap(&tm_a, xh, p.M, Cp, Cp, BM)) p.a_mode = TMA_A_TILED_K;
    else if (!make_im2col_tmap(&tm_a, xh, Cp, W, H, N, pw, ph, fw, fh, sh, sv, BM, false)) return MNV_OK;
    if (b_direct == 1 && make_b_tmap(&tm_b, w, Co, Ci, Ci, (p.wide || p.tall == 2) ? p.bn / 2 : p.bn)) pack_filter = false;
    else if (b_direct == 2 && make_mn3_tmap(&tm_b, w, Co, Ci, Co, p.tall == 2 ? p.bn / 64 : p.bn / 32)) { pack_filter = false; p.b_mn = 2; }
    else if (b_direct == 2) return MNV_OK;     // planned for 32-column boxes: let the caller take its generic route
    else if (!make_b_tmap(&tm_b, wb, Co, p.K, p.K, (p.wide || p.tall == 2) ? p.bn / 2 : p.bn)) return MNV_OK;
  }
  int rc = MNV_OK;
  if (s2d) rc = launch_s2d(x, xh, N, *s2d, Cp, s);
  else if (!(use_twin && twin.valid())) {
    rc = launch_nhwc(x, xh, N, Ci, Cp, H * W, s);
    if (!rc && use_twin) *twin.state = 1;
  }
  if (rc) return rc;
  if (pack_filter) {
    if (s2d) launch_pdl(s2d_filter_pack_kernel, dim3(stream_grid(static_cast<size_t>(Co) * K)), dim3(kBlock), 0, s, w, wb, Co, *s2d, cpt, 0, 4, static_cast<int>(K), w_sn, w_sc, flip, prepass_round());
    else launch_pdl(filter_pack_kernel, dim3(stream_grid(static_cast<size_t>(Co) * K)), dim3(kBlock), 0, s, w, wb, Co, Ci, ff, cpt * BK, w_sn, w_sc, flip, prepass_round());
    rc = finish_launch();
    if (rc) return rc;
  }
  *done = true;
  return launch_umma_tma(transposed ? q : p, tm_a, tm_b, s);
}

// Strided few-channel convolution through its space-to-depth view on the shift-GEMM kernel (see conv_shift_fwd_kernel).
// *done stays false when the geometry does not fit (more than 128 filters, halo above 128 rows, no workspace).
static int conv_shift_fprop(const float* x, const float* w, long long w_sn, long long w_sc, int flip, const float* bias, int relu, float* out,
                            int N, int Co, int Ho, int Wo, const S2D& v, void* ws, size_t ws_bytes, cudaStream_t s, bool* done) {
  *done = false;
  if (g_opt_no_shift.load() || g_opt_simt.load() || g_opt_no_tma.load() || !ws || !get_encode_fn()) return MNV_OK;
  const int halo = (v.fhv - 1) * v.Wv + v.fwv - 1, taps = v.fhv * v.fwv;
  const int cpt = (v.Civ + BK - 1) / BK, Cp = (v.Civ + 3) / 4 * 4, bn = (Co + 15) / 16 * 16;
  if (bn > 128 || halo > kShPlaneRows - kShTileM) return MNV_OK;
  const long long total_pos = static_cast<long long>(N) * v.Hv * v.Wv;
  if (total_pos * Cp >= 0x7fffffffLL) return MNV_OK;
  const size_t x_bytes = round256(static_cast<size_t>(total_pos) * Cp * sizeof(float));
  const int nk_last = (v.Civ - (cpt - 1) * BK + 7) / 8, tps_last = 4 / nk_last;
  const int K = ((cpt - 1) * taps + (taps + tps_last - 1) / tps_last) * BK;   // filter stages x 32
  const size_t b_bytes = round256(static_cast<size_t>(Co) * K * sizeof(float));
  if (ws_bytes < x_bytes + b_bytes) return MNV_OK;
  float* xv = static_cast<float*>(ws);
  float* wb = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + x_bytes);
  ShiftParams p;
  p.bias = bias; p.out = out; p.Co = Co; p.bn = bn; p.Ho = Ho; p.Wo = Wo; p.Hv = v.Hv; p.Wv = v.Wv; p.fwv = v.fwv; p.taps = taps;
  p.cpt = cpt; p.relu = relu; p.total_pos = static_cast<int>(total_pos); p.nk_last = nk_last;
  p.b_stages = (kShSmemMax - kShSmemFixed) / (bn * 128);
  if (p.b_stages > kShBStagesMax) p.b_stages = kShBStagesMax;
  const int smem_bytes = kShSmemFixed + p.b_stages * bn * 128;
  p.m_tiles = static_cast<int>((total_pos + kShTileM - 1) / kShTileM);
  p.plane_boxes = (kShTileM + halo + 127) / 128;
  p.wait_hint = static_cast<unsigned>(g_opt_wait_hint.load());
  p.dbg = g_opt_shift_dbg.load();
  CUtensorMap tm_x, tm_w;
  memset(&tm_x, 0, sizeof(tm_x));
  memset(&tm_w, 0, sizeof(tm_w));
  if (!make_b_tmap(&tm_x, xv, p.total_pos, Cp, Cp, 128) || !make_b_tmap(&tm_w, wb, Co, K, K, bn)) return MNV_OK;
  static std::atomic<uint64_t> attr_done{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_done.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(conv_shift_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kShSmemMax);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  int rc = launch_s2d(x, xv, N, v, Cp, s);
  if (rc) return rc;
  launch_pdl(s2d_filter_pack_kernel, dim3(stream_grid(static_cast<size_t>(Co) * K)), dim3(kBlock), 0, s, w, wb, Co, v, cpt, 1, nk_last, K, w_sn, w_sc, flip, prepass_round());
  rc = finish_launch();
  if (rc) return rc;
  *done = true;
  const int sms = sm_budget();
  launch_pdl(conv_shift_fwd_kernel, dim3(p.m_tiles < sms ? p.m_tiles : sms), dim3(kShThreads), smem_bytes, s, p, tm_x, tm_w);
  return finish_launch();
}

}  // namespace mnv

using namespace mnv;

extern "C" {

#ifdef MNV_TUNING
// Debug / tuning hook (not part of the reference surface; declared in include/mnv_debug.h, tuning build only).  Keys: "simt" (1 routes GEMM/conv through
// the SIMT checker kernel), "max_splits" (clamp split-K), "no_tma" (gather B with threads),
// "no_fwd_bwd" (generic backward-data gather for stride 1).  Returns the previous value, -1 for a bad key.
__attribute__((visibility("default"))) int mnv_debug_set_option(const char* key, int value) {
  if (!key) return -1;
  std::string k(key);
  if (k == "simt") return g_opt_simt.exchange(value);
  if (k == "max_splits") return g_opt_max_splits.exchange(value);
  if (k == "no_tma") return g_opt_no_tma.exchange(value);
  if (k == "no_fwd_bwd") return g_opt_no_fwd_bwd.exchange(value);
  if (k == "no_ktab") return g_opt_no_ktab.exchange(value);
  if (k == "wait_hint") return g_opt_wait_hint.exchange(value);
  if (k == "no_wide") return g_opt_no_wide.exchange(value);
  if (k == "no_tma_a") return g_opt_no_tma_a.exchange(value);
  if (k == "tma_tf32") return g_opt_tma_tf32.exchange(value);
  if (k == "no_deep") return g_opt_no_deep.exchange(value);
  if (k == "sm_budget") return g_opt_sm_budget.exchange(value);
  if (k == "no_klane") return g_opt_no_klane.exchange(value);
  if (k == "no_tall") return g_opt_no_tall.exchange(value);
  if (k == "no_s2d") return g_opt_no_s2d.exchange(value);
  if (k == "pf_dist") return g_opt_pf_dist.exchange(value);
  if (k == "no_tail") return g_opt_no_tail.exchange(value);
  if (k == "no_shift") return g_opt_no_shift.exchange(value);
  if (k == "shift_dbg") return g_opt_shift_dbg.exchange(value);
  if (k == "s2d_im2col") return g_opt_s2d_im2col.exchange(value);
  if (k == "tall_min_stages") return g_opt_tall_min_stages.exchange(value);
  if (k == "force_tma_a") return g_opt_force_tma_a.exchange(value);
  if (k == "no_nhwc_wgrad") return g_opt_no_nhwc_wgrad.exchange(value);
  if (k == "no_mn3") return g_opt_no_mn3.exchange(value);
  if (k == "pair") return g_opt_pair.exchange(value);
  if (k == "no_pointwise") return g_opt_no_pointwise.exchange(value);
  if (k == "wgrad_wide") return g_opt_wgrad_wide.exchange(value);
  if (k == "pair_remote") return g_opt_pair_remote.exchange(value);
  if (k == "no_transposed") return g_opt_no_transposed.exchange(value);
  return -1;
}
#endif

int mnv_matmult(const float* a, const float* b, float* c, int m, int n, int k, void* workspace,
                size_t workspace_bytes, mnv_stream_t stream) {
  if (m < 0 || n < 0 || k < 0) return MNV_EINVAL;
  if (m == 0 || n == 0) return MNV_OK;
  if (!a || !b || !c) return MNV_EINVAL;
  if (k == 0) return mnv_fill(c, static_cast<size_t>(m) * n, 0.f, stream);
  GemmParams p;
  zero_conv(p);
  p.a = a; p.b = b; p.out = c;
  p.M = m; p.N = n; p.K = k;
  p.lda = m; p.ldb = k;
  p.b_vec = (k % 4 == 0) && aligned16(b);
  p.P = m; p.img_stride = 0; p.col_stride = m;
  return launch_gemm<A_COLMAJOR, B_KMAJOR>(p, workspace, workspace_bytes, as_stream(stream));
}

// c{m,n} = op(a) * op(b).  trans_a: a is stored {k,m} (so op(a) is K-major: a plain 2-D tensor map); trans_b: b is
// stored {n,k} (op(b) is MN-major).  Both operands go through TMA as they lie in memory; when an alignment rule of
// the tensor maps is not met the transpose is materialised in the workspace and the plain path runs.
int mnv_matmult_ex(const float* a, const float* b, float* c, int m, int n, int k, int trans_a, int trans_b, void* workspace,
                   size_t workspace_bytes, mnv_stream_t stream) {
  if (!trans_a && !trans_b) return mnv_matmult(a, b, c, m, n, k, workspace, workspace_bytes, stream);
  if (m < 0 || n < 0 || k < 0) return MNV_EINVAL;
  if (m == 0 || n == 0) return MNV_OK;
  if (!a || !b || !c) return MNV_EINVAL;
  if (k == 0) return mnv_fill(c, static_cast<size_t>(m) * n, 0.f, stream);
  cudaStream_t s = as_stream(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  size_t ws_left = workspace ? workspace_bytes : 0;
  // operand-by-operand: can it be mapped in place?
  const bool a_ok = trans_a ? (k % 4 == 0 && aligned16(a)) : (m % 4 == 0 && aligned16(a));
  const bool b_ok = trans_b ? (n % 4 == 0 && aligned16(b)) : (k % 4 == 0 && aligned16(b));
  const bool direct = a_ok && b_ok && !g_opt_simt.load() && !g_opt_no_tma.load() && !(g_opt_no_tma_a.load() & 2) && get_encode_fn() != nullptr;
  if (!direct) {   // materialise the transposes (needs the workspace), then the plain path
    const float* a2 = a;
    const float* b2 = b;
    if (trans_a) {
      size_t need = round256(static_cast<size_t>(m) * k * sizeof(float));
      if (ws_left < need) return MNV_EWORKSPACE;
      int rc = mnv_transpose(a, reinterpret_cast<float*>(ws), k, m, stream);
      if (rc) return rc;
      a2 = reinterpret_cast<float*>(ws); ws += need; ws_left -= need;
    }
    if (trans_b) {
      size_t need = round256(static_cast<size_t>(n) * k * sizeof(float));
      if (ws_left < need) return MNV_EWORKSPACE;
      int rc = mnv_transpose(b, reinterpret_cast<float*>(ws), n, k, stream);
      if (rc) return rc;
      b2 = reinterpret_cast<float*>(ws); ws += need; ws_left -= need;
    }
    return mnv_matmult(a2, b2, c, m, n, k, ws_left ? ws : nullptr, ws_left, stream);
  }
  GemmParams p;
  zero_conv(p);
  p.a = a; p.b = b; p.out = c;
  p.M = m; p.N = n; p.K = k;
  p.lda = trans_a ? k : m; p.ldb = trans_b ? n : k; p.b_vec = 1;
  p.P = m; p.img_stride = 0; p.col_stride = m;
  p.a_mode = trans_a ? TMA_A_TILED_K : TMA_A_TILED_MN;
  p.b_mn = trans_b ? 1 : 0;
  plan_tiles(p, ws_left, !trans_b, true);   // MN-major B: boxes of 32 columns, no wide tile
  if (trans_b) {
    p.bn = p.tall == 2 ? (p.bn + 63) / 64 * 64 : (p.bn + 31) / 32 * 32;   // a CTA of a pair stages bn / 2 columns
    if (p.bn > BN_MAX) p.bn = BN_MAX;
    p.n_tiles = (p.N + p.bn - 1) / p.bn;
    plan_splits(p, ws_left);
  }
  p.partial = p.splits > 1 ? reinterpret_cast<float*>(ws) : nullptr;
  CUtensorMap tm_a, tm_b;
  memset(&tm_a, 0, sizeof(tm_a));
  memset(&tm_b, 0, sizeof(tm_b));
  bool ok_a, ok_b;
  const bool mn3 = !g_opt_no_mn3.load();
  if (trans_a) ok_a = make_b_tmap(&tm_a, a, m, k, k, BM);
  else if (mn3 && make_mn3_tmap(&tm_a, a, m, k, m, p.tall == 1 ? 8 : 4)) { ok_a = true; p.a_g3 = 1; }
  else ok_a = make_a_mn_tmap(&tm_a, a, m, k, m);
  if (!trans_b) ok_b = make_b_tmap(&tm_b, b, n, k, k, (p.wide || p.tall == 2) ? p.bn / 2 : p.bn);
  else if (mn3 && make_mn3_tmap(&tm_b, b, n, k, n, p.tall == 2 ? p.bn / 64 : p.bn / 32)) { ok_b = true; p.b_mn = 2; }
  else ok_b = make_a_mn_tmap(&tm_b, b, n, k, n);
  if (!ok_a || !ok_b) return MNV_EINVAL;
  return launch_umma_tma(p, tm_a, tm_b, s);
}

static int conv_forward_impl(const float* bottom, const float* filter, const float* bias, float* top, int N, int Ci,
                             int Co, int H, int W, int ph, int pw, int sv, int sh, int fh, int fw, void* workspace,
                             size_t workspace_bytes, mnv_stream_t stream, int relu, Twin twin = Twin{nullptr, nullptr});
int mnv_conv_forward(const float* bottom, const float* filter, const float* bias, float* top, int N, int Ci,
                     int Co, int H, int W, int ph, int pw, int sv, int sh, int fh, int fw, void* workspace,
                     size_t workspace_bytes, mnv_stream_t stream) {
  return conv_forward_impl(bottom, filter, bias, top, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw, workspace, workspace_bytes, stream, 0);
}
int mnv_conv_forward_relu(const float* bottom, const float* filter, const float* bias, float* top, int N, int Ci,
                          int Co, int H, int W, int ph, int pw, int sv, int sh, int fh, int fw, void* workspace,
                          size_t workspace_bytes, mnv_stream_t stream) {
  return conv_forward_impl(bottom, filter, bias, top, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw, workspace, workspace_bytes, stream, 1);
}
int mnv_conv_forward_tw(const float* bottom, const float* filter, const float* bias, float* top, int N, int Ci, int Co, int H, int W,
                        int ph, int pw, int sv, int sh, int fh, int fw, int relu, float* bottom_twin, int* bottom_twin_state,
                        void* workspace, size_t workspace_bytes, mnv_stream_t stream) {
  return conv_forward_impl(bottom, filter, bias, top, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw, workspace, workspace_bytes, stream, relu ? 1 : 0,
                           Twin{bottom_twin, bottom_twin_state});
}
int mnv_conv_twin_wanted(int N, int Ci, int Co, int H, int W, int ph, int pw, int sv, int sh, int fh, int fw) {
  if (check_conv(N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw) || N == 0 || !get_im2col_fn() || !get_encode_fn()) return 0;
  const int ff = fh * fw;
  const bool lim = ph <= 127 && pw <= 127 && fh - 1 - ph <= 128 && fw - 1 - pw <= 128 && sv <= 8 && sh <= 8;
  auto chan_ok = [](int c) { return ((c + BK - 1) / BK) * BK * 2 <= c * 3; };
  S2D v;
  const int Ho = (H + 2 * ph - fh) / sv + 1, Wo = (W + 2 * pw - fw) / sh + 1;
  const bool shift = s2d_plan(&v, Ci, H, W, Ho, Wo, ph, pw, sv, sh, fh, fw);
  const bool fwd = !shift && lim && chan_ok(Ci);
  const bool dgrad = sv == 1 && sh == 1 && fh - 1 - ph >= 0 && fw - 1 - pw >= 0 && chan_ok(Co);
  (void)ff;
  const bool wgrad = lim && chan_ok(Ci);
  return ((fwd || wgrad) ? 1 : 0) | ((dgrad || wgrad) ? 2 : 0);
}
int mnv_relu_backward_tw(const float* top, const float* top_diff, float* bottom_diff, int N, int C, int H, int W, float* twin,
                         int* twin_state, mnv_stream_t stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return MNV_EINVAL;
  const size_t n = static_cast<size_t>(N) * C * H * W;
  if (n == 0) return MNV_OK;
  if (!top || !top_diff || !bottom_diff) return MNV_EINVAL;
  const int Cp = (C + 3) / 4 * 4, HW = H * W;
  if (!twin || !twin_state || !fits_int(static_cast<long long>(N) * HW * Cp))    // no twin wanted: the plain op
    return mnv_relu_backward(top, top, top_diff, bottom_diff, N, C, H, W, stream);
  float* tilesum = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(twin) + twin_copy_bytes(N, C, HW));
  int rc = launch_nhwc_t<true>(top_diff, twin, N, C, Cp, HW, as_stream(stream), tilesum, top, bottom_diff);
  if (!rc) *twin_state = 3;     // copy + per-tile channel sums
  return rc;
}
size_t mnv_conv_twin_bytes(int N, int C, int H, int W) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  return twin_copy_bytes(N, C, H * W) + twin_sums_bytes(N, C, H * W);
}
static int conv_forward_impl(const float* bottom, const float* filter, const float* bias, float* top, int N, int Ci,
                             int Co, int H, int W, int ph, int pw, int sv, int sh, int fh, int fw, void* workspace,
                             size_t workspace_bytes, mnv_stream_t stream, int relu, Twin twin) {
  int rc = check_conv(N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw);
  if (rc) return rc;
  if (N == 0) return MNV_OK;
  if (!bottom || !filter || !bias || !top) return MNV_EINVAL;
  {  // strided few-channel convolutions (AlexNet conv1) through their stride-1 space-to-depth view
    S2D v;
    const int Ho = (H + 2 * ph - fh) / sv + 1, Wo = (W + 2 * pw - fw) / sh + 1;
    if (!g_opt_no_s2d.load() && s2d_plan(&v, Ci, H, W, Ho, Wo, ph, pw, sv, sh, fh, fw)) {
      bool done = false;
      rc = conv_shift_fprop(bottom, filter, static_cast<long long>(Ci) * fh * fw, fh * fw, 1, bias, relu, top, N, Co, Ho, Wo, v, workspace,
                            workspace_bytes, as_stream(stream), &done);
      if (rc || done) return rc;
      // the view on the im2col-fed kernel is slower than the gather kernel on AlexNet conv1 (0.48 vs 0.42 ms): experiments only
      if (g_opt_s2d_im2col.load())
      rc = conv_tma_fprop(bottom, filter, static_cast<long long>(Ci) * fh * fw, fh * fw, 1, bias, relu, top, N, v.Civ, Co, v.Hv, v.Wv,
                          Ho, Wo, 0, 0, 1, 1, v.fhv, v.fwv, workspace, workspace_bytes, as_stream(stream), &done, &v);
      if (rc || done) return rc;
    }
  }
  {  // both operands through TMA when the channel count makes 32-channel k-stages worthwhile
    bool done = false;
    rc = conv_tma_fprop(bottom, filter, static_cast<long long>(Ci) * fh * fw, fh * fw, 1, bias, relu, top, N, Ci, Co, H, W,
                        (H + 2 * ph - fh) / sv + 1, (W + 2 * pw - fw) / sh + 1, ph, pw, sv, sh, fh, fw, workspace, workspace_bytes,
                        as_stream(stream), &done, nullptr, twin);
    if (rc || done) return rc;
  }
  GemmParams p;
  zero_conv(p);
  p.Ci = Ci; p.Co = Co; p.H = H; p.W = W; p.fh = fh; p.fw = fw; p.ph = ph; p.pw = pw; p.sv = sv; p.sh = sh;
  p.Ho = (H + 2 * ph - fh) / sv + 1; p.Wo = (W + 2 * pw - fw) / sh + 1;
  long long M = static_cast<long long>(N) * p.Ho * p.Wo, K = static_cast<long long>(Ci) * fh * fw;
  if (!fits_int(M) || !fits_int(K) || !fits_int(static_cast<long long>(N) * Ci * H * W)) return MNV_EUNSUPPORTED;
  p.a = bottom; p.b = filter; p.bias = bias; p.out = top; p.relu = relu;
  p.M = static_cast<int>(M); p.N = Co; p.K = static_cast<int>(K);
  p.ldb = p.K; p.b_vec = (p.K % 4 == 0) && aligned16(filter);
  p.P = p.Ho * p.Wo; p.img_stride = static_cast<long long>(Co) * p.P; p.col_stride = p.P;
  return launch_gemm<A_IM2COL_FWD, B_KMAJOR>(p, workspace, workspace_bytes, as_stream(stream));
}

static int conv_backward_data_impl(const float* top_diff, const float* filter, float* bottom_diff, int N, int Ci, int Co,
                                   int H, int W, int ph, int pw, int sv, int sh, int fh, int fw, void* workspace,
                                   size_t workspace_bytes, mnv_stream_t stream, Twin twin);
int mnv_conv_backward_data(const float* top_diff, const float* filter, float* bottom_diff, int N, int Ci, int Co,
                           int H, int W, int ph, int pw, int sv, int sh, int fh, int fw, void* workspace,
                           size_t workspace_bytes, mnv_stream_t stream) {
  return conv_backward_data_impl(top_diff, filter, bottom_diff, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw, workspace, workspace_bytes, stream,
                                 Twin{nullptr, nullptr});
}
int mnv_conv_backward_data_tw(const float* top_diff, const float* filter, float* bottom_diff, int N, int Ci, int Co, int H, int W,
                              int ph, int pw, int sv, int sh, int fh, int fw, float* top_diff_twin, int* top_diff_twin_state,
                              void* workspace, size_t workspace_bytes, mnv_stream_t stream) {
  return conv_backward_data_impl(top_diff, filter, bottom_diff, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw, workspace, workspace_bytes, stream,
                                 Twin{top_diff_twin, top_diff_twin_state});
}
static int conv_backward_data_impl(const float* top_diff, const float* filter, float* bottom_diff, int N, int Ci, int Co,
                                   int H, int W, int ph, int pw, int sv, int sh, int fh, int fw, void* workspace,
                                   size_t workspace_bytes, mnv_stream_t stream, Twin twin) {
  int rc = check_conv(N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw);
  if (rc) return rc;
  if (N == 0) return MNV_OK;
  if (!top_diff || !filter || !bottom_diff) return MNV_EINVAL;
  // the filter is re-laid out as [ci][(co,r,s)] in the workspace (a few MB at most, once per call)
  size_t wt_bytes = (static_cast<size_t>(Co) * Ci * fh * fw * sizeof(float) + 255) / 256 * 256;
  const int Ho = (H + 2 * ph - fh) / sv + 1, Wo = (W + 2 * pw - fw) / sh + 1;
  // stride 1: backward-data == forward convolution of top_diff with the swapped, 180-degree-rotated
  // filter and pad' = f-1-pad, so it runs on the (fast) forward gather.
  const bool as_forward = sv == 1 && sh == 1 && fh - 1 - ph >= 0 && fw - 1 - pw >= 0 && !g_opt_no_fwd_bwd.load();
  if (!workspace || workspace_bytes < wt_bytes) {
    // scratch-free schedule (mnv.h: a null / short workspace is legal): the producer warps gather B straight out of the
    // filter as stored, [co][ci][r][s] read as B[ci][(co,r,s)]
    GemmParams p;
    zero_conv(p);
    long long M = static_cast<long long>(N) * H * W, K = static_cast<long long>(Co) * fh * fw;
    if (!fits_int(M) || !fits_int(K) || !fits_int(static_cast<long long>(N) * Co * Ho * Wo)) return MNV_EUNSUPPORTED;
    p.a = top_diff; p.b = filter; p.out = bottom_diff;
    p.M = static_cast<int>(M); p.N = Ci; p.K = static_cast<int>(K);
    p.P = H * W; p.img_stride = static_cast<long long>(Ci) * p.P; p.col_stride = p.P;
    p.fh = fh; p.fw = fw; p.b_flip = as_forward ? 1 : 0;
    if (as_forward) {
      p.Ci = Co; p.Co = Ci; p.H = Ho; p.W = Wo; p.Ho = H; p.Wo = W;
      p.ph = fh - 1 - ph; p.pw = fw - 1 - pw; p.sv = 1; p.sh = 1;
      return launch_gemm<A_IM2COL_FWD, B_FILTER_T>(p, nullptr, 0, as_stream(stream));
    }
    p.Ci = Ci; p.Co = Co; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.ph = ph; p.pw = pw; p.sv = sv; p.sh = sh;
    return launch_gemm<A_IM2COL_BWD, B_FILTER_T>(p, nullptr, 0, as_stream(stream));
  }
  float* wt = static_cast<float*>(workspace);
  if (as_forward) {
    // B[n = ci][k = (tap, co)] = filter[co][ci][tap]: the two 180-degree rotations (true convolution, transposed
    // operator) cancel
    bool done = false;
    rc = conv_tma_fprop(top_diff, filter, fh * fw, static_cast<long long>(Ci) * fh * fw, 0, nullptr, 0, bottom_diff, N, Co, Ci, Ho, Wo, H, W,
                        fh - 1 - ph, fw - 1 - pw, 1, 1, fh, fw, workspace, workspace_bytes, as_stream(stream), &done, nullptr, twin);
    if (rc || done) return rc;
  }
  launch_pdl(filter_swap_kernel, dim3(stream_grid(static_cast<size_t>(Co) * Ci * fh * fw)), dim3(kBlock), 0, as_stream(stream), 
      filter, wt, Co, Ci, fh * fw, as_forward ? 1 : 0, prepass_round());
  rc = finish_launch();
  if (rc) return rc;
  GemmParams p;
  zero_conv(p);
  long long M = static_cast<long long>(N) * H * W, K = static_cast<long long>(Co) * fh * fw;
  if (!fits_int(M) || !fits_int(K) || !fits_int(static_cast<long long>(N) * Co * Ho * Wo)) return MNV_EUNSUPPORTED;
  p.a = top_diff; p.b = wt; p.out = bottom_diff;
  p.M = static_cast<int>(M); p.N = Ci; p.K = static_cast<int>(K);
  p.ldb = p.K; p.b_vec = (p.K % 4 == 0);
  p.P = H * W; p.img_stride = static_cast<long long>(Ci) * p.P; p.col_stride = p.P;
  p.fh = fh; p.fw = fw;
  void* ws2 = static_cast<uint8_t*>(workspace) + wt_bytes;
  size_t ws2_bytes = workspace_bytes - wt_bytes;
  if (as_forward) {
    // forward-gather view: source = top_diff (Co channels, Ho x Wo), output pixels = bottom (H x W)
    p.Ci = Co; p.Co = Ci; p.H = Ho; p.W = Wo; p.Ho = H; p.Wo = W;
    p.ph = fh - 1 - ph; p.pw = fw - 1 - pw; p.sv = 1; p.sh = 1;
    return launch_gemm<A_IM2COL_FWD, B_KMAJOR>(p, ws2, ws2_bytes, as_stream(stream));
  }
  p.Ci = Ci; p.Co = Co; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.ph = ph; p.pw = pw; p.sv = sv; p.sh = sh;
  return launch_gemm<A_IM2COL_BWD, B_KMAJOR>(p, ws2, ws2_bytes, as_stream(stream));
}

static int conv_backward_filter_impl(const float* bottom, const float* top_diff, float* filter_diff, int N, int Ci, int Co,
                                     int H, int W, int ph, int pw, int sv, int sh, int fh, int fw, void* workspace,
                                     size_t workspace_bytes, mnv_stream_t stream, float* bias_diff, bool* bias_done,
                                     Twin xtw = Twin{nullptr, nullptr}, Twin dtw = Twin{nullptr, nullptr});
int mnv_conv_backward_filter(const float* bottom, const float* top_diff, float* filter_diff, int N, int Ci, int Co,
                             int H, int W, int ph, int pw, int sv, int sh, int fh, int fw, void* workspace,
                             size_t workspace_bytes, mnv_stream_t stream) {
  bool unused = false;
  return conv_backward_filter_impl(bottom, top_diff, filter_diff, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw, workspace, workspace_bytes,
                                   stream, nullptr, &unused);
}
int mnv_conv_backward_filter_bias(const float* bottom, const float* top_diff, float* filter_diff, float* bias_diff, int N,
                                  int Ci, int Co, int H, int W, int ph, int pw, int sv, int sh, int fh, int fw,
                                  void* workspace, size_t workspace_bytes, mnv_stream_t stream) {
  if (!bias_diff) return MNV_EINVAL;
  int rc = check_conv(N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw);
  if (rc) return rc;
  if (N == 0) {
    rc = mnv_fill(bias_diff, static_cast<size_t>(Co), 0.f, stream);
    if (rc) return rc;
  }
  bool bias_done = false;
  rc = conv_backward_filter_impl(bottom, top_diff, filter_diff, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw, workspace, workspace_bytes,
                                 stream, bias_diff, &bias_done);
  if (rc || bias_done || N == 0) return rc;
  // top_diff needed no re-pitch (16-byte rows) or there was no workspace: the stand-alone reduction
  return mnv_conv_backward_bias(top_diff, bias_diff, N, Co, (H + 2 * ph - fh) / sv + 1, (W + 2 * pw - fw) / sh + 1, workspace,
                                workspace_bytes, stream);
}
int mnv_conv_backward_filter_tw(const float* bottom, const float* top_diff, float* filter_diff, float* bias_diff, int N, int Ci, int Co,
                                int H, int W, int ph, int pw, int sv, int sh, int fh, int fw, float* bottom_twin, int* bottom_twin_state,
                                float* top_diff_twin, int* top_diff_twin_state, void* workspace, size_t workspace_bytes,
                                mnv_stream_t stream) {
  int rc = check_conv(N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw);
  if (rc) return rc;
  if (bias_diff && N == 0) {
    rc = mnv_fill(bias_diff, static_cast<size_t>(Co), 0.f, stream);
    if (rc) return rc;
  }
  bool bias_done = false;
  rc = conv_backward_filter_impl(bottom, top_diff, filter_diff, N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw, workspace, workspace_bytes,
                                 stream, bias_diff, &bias_done, Twin{bottom_twin, bottom_twin_state}, Twin{top_diff_twin, top_diff_twin_state});
  if (rc || !bias_diff || bias_done || N == 0) return rc;
  return mnv_conv_backward_bias(top_diff, bias_diff, N, Co, (H + 2 * ph - fh) / sv + 1, (W + 2 * pw - fw) / sh + 1, workspace,
                                workspace_bytes, stream);
}
static int conv_backward_filter_impl(const float* bottom, const float* top_diff, float* filter_diff, int N, int Ci, int Co,
                                     int H, int W, int ph, int pw, int sv, int sh, int fh, int fw, void* workspace,
                                     size_t workspace_bytes, mnv_stream_t stream, float* bias_diff, bool* bias_done, Twin xtw, Twin dtw) {
  *bias_done = false;
  int rc = check_conv(N, Ci, Co, H, W, ph, pw, sv, sh, fh, fw);
  if (rc) return rc;
  if (!bottom || !top_diff || !filter_diff) return MNV_EINVAL;
  if (N == 0) return mnv_fill(filter_diff, static_cast<size_t>(Co) * Ci * fh * fw, 0.f, stream);
  GemmParams p;
  zero_conv(p);
  p.Ci = Ci; p.Co = Co; p.H = H; p.W = W; p.fh = fh; p.fw = fw; p.ph = ph; p.pw = pw; p.sv = sv; p.sh = sh;
  p.Ho = (H + 2 * ph - fh) / sv + 1; p.Wo = (W + 2 * pw - fw) / sh + 1;
  const int P = p.Ho * p.Wo;
  long long M = static_cast<long long>(Ci) * fh * fw, K = static_cast<long long>(N) * P;
  if (!fits_int(M) || !fits_int(K) || !fits_int(static_cast<long long>(N) * Ci * H * W)) return MNV_EUNSUPPORTED;
  p.a = bottom; p.b = top_diff; p.out = filter_diff;
  p.M = static_cast<int>(M); p.N = Co; p.K = static_cast<int>(K);
  p.P = p.M; p.img_stride = 0; p.col_stride = p.M;  // filter_diff[co][(ci,r,s)]
  cudaStream_t s = as_stream(stream);
  if (g_opt_simt.load() || g_opt_no_tma.load())
    return launch_gemm<A_IM2COL_WGRAD, B_DY_WGRAD>(p, workspace, workspace_bytes, s);

  {  // Both operands channels-last, both MN-major, k = the flat (image, pixel) axis: A = bottom through an im2col map
     // (32 pixels x 4 boxes of 32 channels per tap), B = top_diff[(n, pixel)][co] as a plain 2-D tensor (boxes of 32 co x
     // 32 pixels).  The im2col walk and the rows of B wrap from one image into the next identically, so nothing is padded
     // per image, top_diff needs no re-pitch, and the two copies are the ones forward / backward-data use (twins).
    const int cpt = (Ci + BK - 1) / BK, Cp = (Ci + 3) / 4 * 4, Cop = (Co + 3) / 4 * 4;
    const int tiles_hw = nhwc_tiles_hw(P);
    const bool geom_ok = cpt * BK * 2 <= Ci * 3 && ph <= 127 && pw <= 127 && fh - 1 - ph <= 128 && fw - 1 - pw <= 128 && sv <= 8 && sh <= 8 &&
                         fits_int(static_cast<long long>(N) * H * W * Cp) && fits_int(static_cast<long long>(N) * P * Cop);
    if (geom_ok && !(g_opt_no_tma_a.load() & 4) && !g_opt_no_nhwc_wgrad.load() && get_im2col_fn() && get_encode_fn()) {
      uint8_t* ws = static_cast<uint8_t*>(workspace);
      size_t ws_left = workspace ? workspace_bytes : 0;
      const size_t x_bytes = xtw.usable() ? 0 : round256(static_cast<size_t>(N) * H * W * Cp * sizeof(float));
      const size_t d_bytes = dtw.usable() ? 0 : round256(static_cast<size_t>(N) * P * Cop * sizeof(float));
      const bool ride_bias = bias_diff != nullptr && !dtw.valid();
      const size_t ts_bytes = ride_bias ? round256(static_cast<size_t>(N) * tiles_hw * Co * sizeof(float)) : 0;
      if (ws_left >= x_bytes + d_bytes + ts_bytes) {
        float* xh = xtw.usable() ? xtw.ptr : reinterpret_cast<float*>(ws);
        float* dyh = dtw.usable() ? dtw.ptr : reinterpret_cast<float*>(ws + x_bytes);
        float* tilesum = ride_bias ? reinterpret_cast<float*>(ws + x_bytes + d_bytes) : nullptr;
        ws += x_bytes + d_bytes + ts_bytes; ws_left -= x_bytes + d_bytes + ts_bytes;
        GemmParams q = p;
        q.a = xh; q.b = dyh; q.M = fh * fw * cpt * BK; q.a_mode = TMA_A_IM2COL_MN; q.b_mn = 1; q.cpt = cpt; q.out_mode = 1; q.spi = 0;
        q.P = q.M; q.col_stride = static_cast<long long>(Ci) * fh * fw; q.ldb = Cop; q.b_vec = 1;
        // More than 256 filters (conv3 / conv4: 384): the wide 128 x 384 tile (two UMMA halves share the A tile, B = one 3-D box
        // of 12 slabs) issues 4 im2col boxes per 128 x 384 x 32 of MMA work where the tall 256 x 192 tile issues 8 -- this path
        // is bound by its A boxes: conv3 backward-filter 0.175 -> 0.137 ms, conv4 0.268 -> 0.210.
        const bool try_wide = g_opt_wgrad_wide.load() && !g_opt_no_mn3.load() && Co > BN_MAX && Cop % 32 == 0;
        plan_tiles(q, ws_left, try_wide, !try_wide);
        if (try_wide && !q.wide) plan_tiles(q, ws_left, false, true);     // mainloop too short for the single-accumulator wide tile
        q.bn = q.tall == 2 ? (q.bn + 63) / 64 * 64 : (q.bn + 31) / 32 * 32;   // MN-major B: boxes of 32 columns (per CTA of a pair)
        if (q.bn > BN_MAX && !q.wide) q.bn = BN_MAX;
        q.n_tiles = (q.N + q.bn - 1) / q.bn;
        plan_splits(q, ws_left);
        q.partial = q.splits > 1 ? reinterpret_cast<float*>(ws) : nullptr;
        CUtensorMap tm_a, tm_b;
        memset(&tm_a, 0, sizeof(tm_a));
        memset(&tm_b, 0, sizeof(tm_b));
        bool ok_b = !g_opt_no_mn3.load() && make_mn3_tmap(&tm_b, dyh, Co, q.K, Cop, q.tall == 2 ? q.bn / 64 : q.bn / 32);
        if (ok_b) q.b_mn = 2;
        else ok_b = make_a_mn_tmap(&tm_b, dyh, Co, q.K, Cop);
        // 1x1 / stride 1 / no padding: the channels-last bottom is the MN-major A matrix [pixel][Cp] itself -- one 3-D box per
        // k-stage instead of four (tall: eight) im2col boxes
        bool ok_a = false;
        if (ok_b && fh == 1 && fw == 1 && sv == 1 && sh == 1 && ph == 0 && pw == 0 && !g_opt_no_pointwise.load() && !g_opt_no_mn3.load() &&
            make_mn3_tmap(&tm_a, xh, Ci, q.K, Cp, q.tall == 1 ? 8 : 4)) {
          ok_a = true; q.a_mode = TMA_A_TILED_MN; q.a_g3 = 1;
        }
        if (!ok_a) ok_a = ok_b && make_im2col_tmap(&tm_a, xh, Cp, W, H, N, pw, ph, fw, fh, sh, sv, BK, true);
        if (ok_a) {
          if (!xtw.valid()) {
            rc = launch_nhwc(bottom, xh, N, Ci, Cp, H * W, s);
            if (rc) return rc;
            if (xtw.usable()) *xtw.state = 1;
          }
          if (!dtw.valid()) {
            rc = launch_nhwc(top_diff, dyh, N, Co, Cop, P, s, tilesum);
            if (rc) return rc;
            if (dtw.usable()) *dtw.state = 1;
            if (tilesum) {   // ConvBackwardBias rides on the pass: per-(image, pixel tile, channel) sums, folded in a fixed order
              launch_pdl(rowsum_fold_kernel, dim3((Co + 31) / 32), dim3(1024), 0, s, tilesum, bias_diff, N * tiles_hw, Co);
              rc = finish_launch();
              if (rc) return rc;
              *bias_done = true;
            }
          }
          if (bias_diff && !*bias_done && dtw.has_sums()) {   // the twin's filler (mnv_relu_backward_tw) left the per-tile sums
            const float* sums = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(dtw.ptr) + twin_copy_bytes(N, Co, P));
            launch_pdl(rowsum_fold_kernel, dim3((Co + 31) / 32), dim3(1024), 0, s, sums, bias_diff, N * tiles_hw, Co);
            rc = finish_launch();
            if (rc) return rc;
            *bias_done = true;
          }
          return launch_umma_tma(q, tm_a, tm_b, s);
        }
      }
    }
  }
  // TMA-fed top_diff: rows must be 16-byte pitched.  When Ho*Wo % 4 != 0 the tensor is re-pitched
  // into the workspace first (one streaming pass, << the GEMM).  Each image's pixel range is padded to
  // a whole number of 32-pixel k-stages so a TMA box never straddles two images.
  const int pitch = (P + 3) / 4 * 4;
  const float* dy_tma = top_diff;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  size_t ws_left = workspace ? workspace_bytes : 0;
  if (pitch != P || !aligned16(top_diff)) {
    size_t need = (static_cast<size_t>(N) * Co * pitch * sizeof(float) + 255) / 256 * 256;
    if (ws_left < need) return launch_gemm<A_IM2COL_WGRAD, B_DY_WGRAD>(p, workspace, workspace_bytes, s);
    float* packed = reinterpret_cast<float*>(ws);
    size_t rows = static_cast<size_t>(N) * Co;
    ws += need; ws_left -= need;
    // the bias gradient rides on this pass: per-(image, channel) sums, folded over the images afterwards
    float* rowsum = nullptr;
    const size_t rs_bytes = round256(rows * sizeof(float));
    if (bias_diff && ws_left >= rs_bytes) { rowsum = reinterpret_cast<float*>(ws); ws += rs_bytes; ws_left -= rs_bytes; }
    launch_pdl(repitch_kernel, dim3(stream_grid(rows * 32)), dim3(kBlock), 0, s, top_diff, packed, P, pitch, rows, prepass_round(), rowsum);
    rc = finish_launch();
    if (rc) return rc;
    if (rowsum) {
      launch_pdl(rowsum_fold_kernel, dim3((Co + 31) / 32), dim3(1024), 0, s, rowsum, bias_diff, N, Co);
      rc = finish_launch();
      if (rc) return rc;
      *bias_done = true;
    }
    dy_tma = packed;
  }
  p.spi = (P + BK - 1) / BK;
  long long kpad = static_cast<long long>(N) * p.spi * BK;
  if (!fits_int(kpad)) return MNV_EUNSUPPORTED;
  p.K = static_cast<int>(kpad);          // k-stages = N * spi; validity is per-pixel inside the gathers
  {  // strided few-channel convolutions: the same all-TMA kernel over the stride-1 space-to-depth view
    S2D v;
    if (!(g_opt_no_tma_a.load() & 4) && !g_opt_no_s2d.load() && g_opt_s2d_im2col.load() && get_im2col_fn() && s2d_plan(&v, Ci, H, W, p.Ho, p.Wo, ph, pw, sv, sh, fh, fw)) {
      const int cpt = (v.Civ + BK - 1) / BK, Cp = (v.Civ + 3) / 4 * 4;
      const size_t x_bytes = round256(static_cast<size_t>(N) * v.Hv * v.Wv * Cp * sizeof(float));
      if (ws_left >= x_bytes && fits_int(static_cast<long long>(N) * v.Hv * v.Wv * Cp)) {
        GemmParams q = p;
        float* xh = reinterpret_cast<float*>(ws);
        q.Ci = v.Civ; q.H = v.Hv; q.W = v.Wv; q.fh = v.fhv; q.fw = v.fwv; q.ph = q.pw = 0; q.sv = q.sh = 1;
        q.r_ci = Ci; q.r_fh = fh; q.r_fw = fw; q.r_sv = sv; q.r_sh = sh;
        q.a = xh; q.M = v.fhv * v.fwv * cpt * BK; q.a_mode = TMA_A_IM2COL_MN; q.cpt = cpt; q.out_mode = 2;
        q.P = q.M; q.col_stride = static_cast<long long>(Ci) * fh * fw;
        plan_tiles(q, ws_left - x_bytes, true, true);
        q.partial = q.splits > 1 ? reinterpret_cast<float*>(ws + x_bytes) : nullptr;
        CUtensorMap tm_a, tm_b;
        memset(&tm_a, 0, sizeof(tm_a));
        memset(&tm_b, 0, sizeof(tm_b));
        if (make_im2col_tmap(&tm_a, xh, Cp, v.Wv, v.Hv, N, 0, 0, v.fwv, v.fhv, 1, 1, BK, true) &&
            make_dy_tmap(&tm_b, dy_tma, P, pitch, Co, N, (q.wide || q.tall == 2) ? q.bn / 2 : q.bn)) {
          rc = launch_s2d(bottom, xh, N, v, Cp, s);
          if (rc) return rc;
          return launch_umma_tma(q, tm_a, tm_b, s);
        }
      }
    }
  }
  {  // A through TMA as well: channels-last copy of the bottom read by an im2col map, pixels as the k axis
    const int cpt = (Ci + BK - 1) / BK, Cp = (Ci + 3) / 4 * 4;
    const size_t x_bytes = round256(static_cast<size_t>(N) * H * W * Cp * sizeof(float));
    if (!(g_opt_no_tma_a.load() & 4) && cpt * BK * 2 <= Ci * 3 && ph <= 127 && pw <= 127 && fh - 1 - ph <= 128 && fw - 1 - pw <= 128 &&
        sv <= 8 && sh <= 8 && get_im2col_fn() && ws_left >= x_bytes && fits_int(static_cast<long long>(N) * H * W * Cp)) {
      GemmParams q = p;
      float* xh = reinterpret_cast<float*>(ws);
      q.a = xh; q.M = fh * fw * cpt * BK; q.a_mode = TMA_A_IM2COL_MN; q.cpt = cpt; q.out_mode = 1;
      q.P = q.M; q.col_stride = static_cast<long long>(Ci) * fh * fw;
      plan_tiles(q, ws_left - x_bytes, true, true);
      q.partial = q.splits > 1 ? reinterpret_cast<float*>(ws + x_bytes) : nullptr;
      CUtensorMap tm_a, tm_b;
      memset(&tm_a, 0, sizeof(tm_a));
      memset(&tm_b, 0, sizeof(tm_b));
      if (make_im2col_tmap(&tm_a, xh, Cp, W, H, N, pw, ph, fw, fh, sh, sv, BK, true) &&
          make_dy_tmap(&tm_b, dy_tma, P, pitch, Co, N, (q.wide || q.tall == 2) ? q.bn / 2 : q.bn)) {
        rc = launch_nhwc(bottom, xh, N, Ci, Cp, H * W, s);
        if (rc) return rc;
        return launch_umma_tma(q, tm_a, tm_b, s);
      }
    }
  }
  plan_tiles(p, ws_left, true);
  p.partial = p.splits > 1 ? reinterpret_cast<float*>(ws) : nullptr;
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  if (!make_dy_tmap(&tm, dy_tma, P, pitch, Co, N, (p.wide || p.tall == 2) ? p.bn / 2 : p.bn)) {
    p.spi = 0; p.K = static_cast<int>(K);
    return launch_gemm<A_IM2COL_WGRAD, B_DY_WGRAD>(p, ws, ws_left, s);
  }
  if ((sv > 1 || sh > 1) && !g_opt_no_klane.load() && !p.wide) rc = launch_umma<A_IM2COL_WGRAD_M, B_KMAJOR, true>(p, tm, s);
  else rc = launch_umma<A_IM2COL_WGRAD, B_KMAJOR, true>(p, tm, s);
  if (rc || p.splits == 1) return rc;
  launch_pdl(splitk_reduce_kernel, dim3(stream_grid(static_cast<size_t>(p.M) * p.N)), dim3(kBlock), 0, s, p);
  return finish_launch();
}

}  // extern "C"
