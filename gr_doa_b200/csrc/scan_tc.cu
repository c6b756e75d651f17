// scan_tc.cu -- stage 2b + 4 on the tensor cores: the pseudo-spectrum scan as a batched steering x noise-subspace contraction
// (tcgen05 kind::tf32, 3xTF32 split), peak picking straight out of TMEM, refinement with the reference arithmetic.
//
// Replaces the angle loop of MUSIC_lin_array_impl::work (gr-doa lib/MUSIC_lin_array_impl.cc:137-142) and find_local_max
// (lib/find_local_max_impl.cc:80-165,186-188) for arrays of up to 16 elements.
//
// For a ULA the null spectrum is the real trigonometric polynomial (scan.cu)
//     Q_f(i) = u_0 + 2 Re sum_{l>=1} u_l z_i^l = sum_k A[f][k] B[i][k],
//     A[f] = (u_0, 2 Re u_1, -2 Im u_1, 2 Re u_2, -2 Im u_2, ...)     B[i] = (1, cos psi_i, sin psi_i, cos 2 psi_i, sin 2 psi_i, ...)
// i.e. ONE real GEMM [frames x 2M-1] x [2M-1 x P] per batch: 128 frames on the MMA's M axis, 128 bins on its N axis, K = 2M-1
// padded to a multiple of 8.  The Horner scan spent 4 (M-1) FMAs per frame and bin on the FP32 pipe (the pipe the covariance
// and the eigensolver are bound by); here the contraction runs on the tensor pipe and the CUDA cores only compare.  fp32
// accuracy comes from the split x = hi + lo (hi = x rounded to tf32):  A_hi B_hi + A_hi B_lo + A_lo B_hi.
//
// B is the constant steering-power table of the plan, split on the host from float64 and stored in global memory as the
// exact shared-memory images of its bin tiles (K-major, SWIZZLE_128B: one 128-byte row of 32 k-values per bin), so a tile is
// one bulk async copy.  A is written by the frames' own threads straight into TMEM (lane = frame, column = k).
// TMEM lane = frame means a thread reads the consecutive bins of ITS frame from the accumulator columns.
//
// Peak detection is STATELESS per bin, so that two warp sets can take alternate bin tiles: with (a, b, c, d) the values of
// bins i-1 .. i+2, bin i is a local minimum of Q (a peak of the spectrum, find_local_max_impl.cc:89-114: entered by a strict
// move, left by a strict move, flats inherit the next strict move) iff  b < a and (c > b or (c == b and d > b));  a flat of
// three or more equal values after a descent (b < a, b == c == d, not running into the end of the vector) is not decided
// locally: the thread walks to the end of the flat (its values of the chunk parked in shared memory, carried over to the next
// chunk of the tile if need be); only a flat that crosses into the next tile flags the frame, which is then redone by the exact
// sequential walker (scan_frame_peaks) -- about one frame in a thousand.
// The test needs one bin to the left and two to the right, so consecutive tiles overlap by three bins: tile t holds bins
// 125 t - 1 .. 125 t + 126 and decides bins 125 t .. 125 t + 124.  The cheap pre-test  b <= min(a, c)  over groups of eight bins
// keeps the detail code (and the candidate-list inserts) off the common path.
//
// One CTA per SM, 12 warps, tiles of 128 frames:
//   warps 0..3   frame threads, even bin tiles (accumulator 0); they also write the A operand
//   warps 4..7   frame threads, odd bin tiles (accumulator 1)
//   warp  8      MMA issuer: per bin tile 3 x ksteps UMMAs (M128 N128 K8), commit
//   warp  9      table loader: 32 KB bulk copies into a 4-stage ring
//   all warps    afterwards: merge of the two candidate lists, refinement of the picked bins with the reference's v^H G v
//                arithmetic (scan_device.cuh; the frame's projector staged in shared memory), outputs
// TMEM: accumulators @0 and @128, A @256 (32 hi | 32 lo columns).
#include "scan_device.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace doa {
namespace {

constexpr int ST_FRAMES = 128;             // frames per tile = MMA M
constexpr int ST_BINS = 128;               // bins per MMA = MMA N
constexpr int ST_STRIDE = 125;             // bins a tile decides (it needs one bin to the left and two to the right)
constexpr int ST_KMAX = 32;                // k-values per table row (one 128-byte swizzle atom)
constexpr int ST_STAGES = 3;               // table ring depth
constexpr int ST_TILE_BYTES = ST_BINS * 128;      // one image (hi or lo) of a bin tile
constexpr int ST_STAGE_BYTES = 2 * ST_TILE_BYTES; // hi image | lo image
constexpr int ST_WARPS = 12;
constexpr int ST_THREADS = ST_WARPS * 32;
constexpr int ST_MMA_WARP = 8, ST_LOAD_WARP = 9;
constexpr uint32_t ST_A_COL = 256;
constexpr int ST_KL = 4;                   // candidate list length (K <= 4)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(b)) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (herk_tc.cu): start >> 4, LBO = 16 B, SBO = 1024 B, version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                  "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                  "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                  "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
  uint32_t v[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
               "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                 "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                 "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                 "=r"(v[31]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}


// Per-frame results of the two detector sets, in shared memory (structure of arrays: index [..][frame], conflict-free).
struct TileResults {
  float cval[2][ST_KL][ST_FRAMES];   // candidate values, best first
  int cbin[2][ST_KL][ST_FRAMES];     // their bins (0x7fffffff = empty)
  int cord[2][ST_KL][ST_FRAMES];     // ordinal of the candidate among its bin tile's peaks (1-based)
  int nem[2][ST_FRAMES];             // local minima found by this set
  int flag[2][ST_FRAMES];            // a flat the local rule cannot decide: redo the frame with the sequential walker
  float q0[ST_FRAMES], qe[ST_FRAMES];   // coarse values of bins 0 and P - 1
  float gv[2][ST_FRAMES]; int gi[2][ST_FRAMES];   // K == 1: first arg-min per set
};

template <int MT, bool ARGMAX>
__global__ void __launch_bounds__(ST_THREADS, 1)
scan_tc_kernel(const float2* __restrict__ u, const float2* __restrict__ G, const uint8_t* __restrict__ tctab, int ksteps,
               const float2* __restrict__ zplain, const float* __restrict__ zpair, const float2* __restrict__ Vtab,
               const float* __restrict__ xaxis, int M, int P, int nframes, int K, float* __restrict__ out_val,
               float* __restrict__ out_loc, int* __restrict__ out_bin, int dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset, not an integer round trip (herk_tc.cu)
  const int nbt = (P + ST_STRIDE - 1) / ST_STRIDE;                               // bin tiles per frame tile
  uint8_t* ring = smem;                                                          // [ST_STAGES][hi image | lo image]
  TileResults* res = reinterpret_cast<TileResults*>(smem + ST_STAGES * ST_STAGE_BYTES);
  float2* gstage = reinterpret_cast<float2*>(res + 1);                            // [ST_WARPS][2][M*M] projector staging (refinement)
  float2* us_all = gstage + (size_t)ST_WARPS * 2 * M * M;                         // [ST_WARPS][M] (fallback paths only)
  float* park = reinterpret_cast<float*>(us_all + ST_WARPS * M);                  // [34][8 warps * 32] a chunk's values per frame thread
  unsigned char* cnt = reinterpret_cast<unsigned char*>(park + 34 * 8 * 32);      // [nbt][ST_FRAMES] peaks per bin tile
  __shared__ uint64_t b_full[ST_STAGES], b_empty[ST_STAGES], d_full[2], d_empty[2], a_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < ST_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&d_full[b], 1); mbar_init(&d_empty[b], ST_FRAMES); }
    mbar_init(&a_full, ST_FRAMES);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == ST_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;

  const int ntiles = (nframes + ST_FRAMES - 1) / ST_FRAMES;
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int uses0 = (nbt + 1) / 2, uses1 = nbt / 2;    // bin tiles per frame tile on accumulator 0 (even) / 1 (odd)
  // The last bin tile is shifted left so that bin P - 1 always sits in its column 125 (it then overlaps its predecessor by more
  // than three bins; it still only decides the bins from 125 (nbt - 1) on).  P >= 126, so there are at least two tiles.

  for (int mt = 0; mt < my_tiles; ++mt) {
    const long long f0 = ((long long)blockIdx.x + (long long)mt * gridDim.x) * ST_FRAMES;
    const long long it = (long long)mt * nbt;            // bin tiles this CTA has been through (ring position)
    if (warp < 8) {
      // ================================ frame threads ================================
      const int set = warp >> 2;                           // 0: even bin tiles, 1: odd bin tiles
      const int fl = (warp & 3) * 32 + lane;               // frame within the tile = TMEM lane
      const long long f = f0 + fl;
      const uint32_t lane_addr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
      if (set == 0) {
        // A row: (u0, 2 Re u1, -2 Im u1, ...), zero beyond 2M - 1 and for frames past the batch
        float a[ST_KMAX];
#pragma unroll
        for (int k = 0; k < ST_KMAX; ++k) a[k] = 0.0f;
        if (f < nframes) {
          const float2* uf = u + f * M;
          a[0] = uf[0].x;
#pragma unroll
          for (int l = 1; l < 16; ++l)
            if (l < M) { const float2 c = uf[l]; a[2 * l - 1] = 2.0f * c.x; a[2 * l] = -2.0f * c.y; }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { hi[j] = to_tf32(a[16 * h + j]); lo[j] = a[16 * h + j] - hi[j]; }
          tmem_st16(lane_addr + ST_A_COL + (uint32_t)h * 16u, hi);
          tmem_st16(lane_addr + ST_A_COL + 32u + (uint32_t)h * 16u, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&a_full);
      }
      CandList<ST_KL, false> list; list.init();
      int n_emit = 0; bool inexact = false;
      float gv = INFINITY; int gi = 0x7fffffff;
      float q_first = 0.0f, q_last = 0.0f;
      const long long use_base = (long long)mt * (set ? uses1 : uses0);
      int use = 0;
      for (int n = set; n < nbt; n += 2, ++use) {
        mbar_wait(&d_full[set], (uint32_t)(use_base + use) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int tile_bin0 = (n == nbt - 1) ? P - 126 : n * ST_STRIDE - 1;   // bin of accumulator column 0
        const int lo_valid = max(1, n * ST_STRIDE);          // bins below belong to the previous tile; bin 0 is never a peak
        int n_tile = 0, pend_bin = -1; float pend_val = 0.f;
        float w0 = 0.f, w1 = 0.f, w2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < ST_BINS / 32; ++c) {
          float W[35];
          {
            float v[32];
            tmem_ld32(lane_addr + (uint32_t)set * 128u + (uint32_t)c * 32u, v);
            if (c == 0) { w0 = w1 = w2 = v[0]; }
            W[0] = w0; W[1] = w1; W[2] = w2;
#pragma unroll
            for (int j = 0; j < 32; ++j) W[3 + j] = v[j];
            w0 = v[29]; w1 = v[30]; w2 = v[31];
          }
          const int base = tile_bin0 + c * 32 - 3;             // W[p] is bin base + p
          if constexpr (ARGMAX) {
            // first arg-min over the bins this tile decides (columns 1 .. 125 = W[p], p + 32 c - 3 in [1, 125])
            float m = INFINITY;
#pragma unroll
            for (int p = 3; p < 35; ++p) m = fminf(m, W[p]);
            if (m <= gv) {
#pragma unroll
              for (int p = 34; p >= 3; --p) {
                const int bin = min(max(base + p, 0), P - 1);
                if (W[p] < gv || (W[p] == gv && bin < gi)) { gv = W[p]; gi = bin; }
              }
            }
          } else {
            if (n == 0 && c == 0) q_first = W[4];              // column 1 of tile 0 = bin 0
            if (n == nbt - 1 && c == 3) q_last = W[3 + 29];       // column 125 of the last tile = bin P - 1
            // cheap pre-test per group of eight bins, then the exact local rule as two bit masks (straight-line code, no inserts)
            unsigned pkmask = 0u, ixmask = 0u;
#pragma unroll
            for (int s8 = 0; s8 < 4; ++s8) {
              bool trig = false;
#pragma unroll
              for (int p = 8 * s8 + 1; p <= 8 * s8 + 8; ++p) trig = trig | (W[p] <= fminf(W[p - 1], W[p + 1]));
              if (trig) {
#pragma unroll
                for (int p = 8 * s8 + 1; p <= 8 * s8 + 8; ++p) {
                  const float a = W[p - 1], b = W[p], cc = W[p + 1], d = W[p + 2];
                  const bool desc = b < a;
                  pkmask |= (desc && (cc > b || (cc == b && d > b))) ? (1u << (p - 1)) : 0u;
                  ixmask |= (desc && cc == b && d == b) ? (1u << (p - 1)) : 0u;
                }
              }
            }
            if ((pkmask | ixmask) != 0u || pend_bin >= 0) {
              // positions p = 1 .. 32 this tile decides: lo_valid <= base + p <= P - 2; a flat running into the last bin is not a peak
              const int plo = max(1, lo_valid - base), phi = min(32, P - 2 - base), pix = min(32, P - 3 - base);
              const unsigned vmask = (phi >= plo) ? ((0xffffffffu >> (32 - phi)) & (0xffffffffu << (plo - 1))) : 0u;
              const unsigned xmask = (pix >= plo) ? ((0xffffffffu >> (32 - pix)) & (0xffffffffu << (plo - 1))) : 0u;
              pkmask &= vmask;
              ixmask &= xmask;
              if ((pkmask | ixmask) != 0u || pend_bin >= 0) {
                // park the chunk's values in this thread's shared-memory column: the loops below index them at run time
                float* col = park + (warp * 32 + lane);
                constexpr int CS = 8 * 32;                       // column stride: W[p] at col[(p - 1) * CS]
#pragma unroll
                for (int p = 1; p <= 34; ++p) col[(p - 1) * CS] = W[p];
                if (pend_bin >= 0) {
                  // a flat of three or more equal values from the previous chunk: the first value not yet seen is W[3]
                  int j = 3;
                  while (j <= 34 && col[(j - 1) * CS] == pend_val) ++j;
                  if (j <= 34) {
                    if (col[(j - 1) * CS] > pend_val) { ++n_tile; if (pend_val < list.val[ST_KL - 1]) list.insert(pend_val, pend_bin, n_tile); }
                    pend_bin = -1;
                  }
                }
                unsigned both = pkmask | ixmask;
                while (both) {
                  const int p0 = __ffs(both) - 1;                // position p = p0 + 1
                  both &= both - 1u;
                  const float b = col[p0 * CS];
                  bool is_pk = ((pkmask >> p0) & 1u) != 0u;
                  if (!is_pk) {
                    // b < a, b == W[p + 1] == W[p + 2]: walk to the end of the flat; it is a peak iff it is left upwards
                    int j = p0 + 4;
                    while (j <= 34 && col[(j - 1) * CS] == b) ++j;
                    if (j > 34) { pend_bin = base + p0 + 1; pend_val = b; continue; }
                    is_pk = col[(j - 1) * CS] > b;
                  }
                  if (is_pk) { ++n_tile; if (b < list.val[ST_KL - 1]) list.insert(b, base + p0 + 1, n_tile); }
                }
              }
            }
          }
        }
        if (pend_bin >= 0) {
          // the flat runs past this tile: past the end of the vector it is no peak (the trailing-flat rule); anywhere else the
          // frame is redone by the sequential walker
          if (n != nbt - 1) inexact = true;
          pend_bin = -1;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&d_empty[set]);
        if constexpr (!ARGMAX) { cnt[n * ST_FRAMES + fl] = (unsigned char)n_tile; n_emit += n_tile; }
      }
      if constexpr (ARGMAX) {
        res->gv[set][fl] = gv; res->gi[set][fl] = gi;
      } else {
#pragma unroll
        for (int k = 0; k < ST_KL; ++k) { res->cval[set][k][fl] = list.val[k]; res->cbin[set][k][fl] = list.idx[k]; res->cord[set][k][fl] = list.ord[k]; }
        res->nem[set][fl] = n_emit;
        res->flag[set][fl] = inexact ? 1 : 0;
        if (set == 0) res->q0[fl] = q_first;
        if (set == ((nbt - 1) & 1)) res->qe[fl] = q_last;
      }
    } else if (warp == ST_MMA_WARP) {
      // ================================ MMA issuer ================================
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ST_BINS >> 3) << 17) | ((uint32_t)(ST_FRAMES >> 4) << 24);
      const uint64_t desc0 = umma_desc(smem_u32(ring));
      const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo0 = (uint32_t)desc0;
      mbar_wait(&a_full, (uint32_t)mt & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int n = 0; n < nbt; ++n) {
        const long long i2 = it + n;
        const int s = (int)(i2 % ST_STAGES), b = n & 1;
        const long long useb = (long long)mt * (b ? uses1 : uses0) + (n >> 1);   // how often accumulator b has been used before
        mbar_wait(&b_full[s], (uint32_t)(i2 / ST_STAGES) & 1u);
        if (useb >= 1) mbar_wait(&d_empty[b], (uint32_t)(useb + 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t bh = desc_lo0 + (uint32_t)s * (ST_STAGE_BYTES >> 4), bl = bh + (ST_TILE_BYTES >> 4);
          const uint32_t dacc = tmem_d + (uint32_t)b * 128u, ahi = tmem_d + ST_A_COL, alo = ahi + 32u;
          for (int k = 0; k < ksteps; ++k)      // the small terms first, the hi * hi term last
            umma_tf32_ts(dacc, alo + k * 8, ((uint64_t)desc_hi << 32) | (bh + (uint32_t)k * 2u), idesc, k == 0 ? 0u : 1u);
          for (int k = 0; k < ksteps; ++k)
            umma_tf32_ts(dacc, ahi + k * 8, ((uint64_t)desc_hi << 32) | (bl + (uint32_t)k * 2u), idesc, 1u);
          for (int k = 0; k < ksteps; ++k)
            umma_tf32_ts(dacc, ahi + k * 8, ((uint64_t)desc_hi << 32) | (bh + (uint32_t)k * 2u), idesc, 1u);
          umma_commit(&b_empty[s]);
          umma_commit(&d_full[b]);
        }
        __syncwarp();
      }
    } else if (warp == ST_LOAD_WARP) {
      // ================================ table loader ================================
      if (lane == 0) {
        for (int n = 0; n < nbt; ++n) {
          const long long i2 = it + n;
          const int s = (int)(i2 % ST_STAGES);
          if (i2 >= ST_STAGES) mbar_wait(&b_empty[s], (uint32_t)(i2 / ST_STAGES + 1) & 1u);
          mbar_expect_tx(&b_full[s], ST_STAGE_BYTES);
          bulk_g2s(ring + (size_t)s * ST_STAGE_BYTES, tctab + (size_t)n * ST_STAGE_BYTES, ST_STAGE_BYTES, &b_full[s]);
        }
      }
      __syncwarp();
    }
    __syncthreads();                          // the tile's frame results are in shared memory
    // ================================ refinement + outputs: a warp per frame ================================
    const int nt = (int)min((long long)ST_FRAMES, (long long)nframes - f0);
    const int MM = M * M;
    float2* us = us_all + warp * M;
    float2* gs = gstage + (size_t)warp * 2 * MM;
    // the frame's projector goes through registers into this warp's staging buffer, one frame ahead of its use
    constexpr int GREG = (MT > 0 ? MT * MT : 256) / 64 > 0 ? (MT > 0 ? MT * MT : 256) / 64 : 1;   // float4 per lane
    float4 gnext[GREG];
    auto g_fetch = [&](int i) {
      const float4* src = reinterpret_cast<const float4*>(G + (f0 + i) * (long long)MM);
#pragma unroll
      for (int q = 0; q < GREG; ++q) { const int e = q * 32 + lane; gnext[q] = (2 * e < MM) ? src[e] : make_float4(0.f, 0.f, 0.f, 0.f); }
    };
    auto g_park = [&](int buf) {
      float4* dst = reinterpret_cast<float4*>(gs + (size_t)buf * MM);
#pragma unroll
      for (int q = 0; q < GREG; ++q) { const int e = q * 32 + lane; if (2 * e < MM) dst[e] = gnext[q]; }
    };
    const int i_first = warp, i_end = (dbg & 1) ? 0 : nt;
    if (i_first < i_end) { g_fetch(i_first); g_park(0); }
    int buf = 0;
    for (int i = i_first; i < i_end; i += ST_WARPS, buf ^= 1) {
      const long long f = f0 + i;
      const bool more = i + ST_WARPS < i_end;
      if (more) g_fetch(i + ST_WARPS);
      __syncwarp();
      const float2* Gf = gs + (size_t)buf * MM;
      float* ov = out_val + f * K; float* ol = out_loc + f * K; int* ob = out_bin ? out_bin + f * K : nullptr;
      if ((dbg & 4) && !ARGMAX) {
        if (lane == 0 && ob) { ob[0] = res->flag[0][i]; ob[1] = res->flag[1][i]; }
      } else if constexpr (ARGMAX) {
        const float v0 = res->gv[0][i], v1 = res->gv[1][i]; const int i0 = res->gi[0][i], i1 = res->gi[1][i];
        const int bi = (v1 < v0 || (v1 == v0 && i1 < i0)) ? i1 : i0;
        argmax_refine_emit<MT, true>(bi, Gf, Vtab, xaxis, M, P, lane, ov, ol, ob);
      } else if ((res->flag[0][i] | res->flag[1][i]) && !(dbg & 2)) {
        // a long flat: the exact sequential walker on the Horner form (runtime-M path, z table read in place)
        scan_frame_peaks<MT, ST_KL>(u + f * M, G + f * (long long)MM, ztab_view(zpair, P), us, Vtab, xaxis, M, P, K, lane, ov, ol, ob);
      } else {
        // merge the two sets' lists: lane l < 8 holds entry (set l / 4, position l % 4); rank by (value, bin)
        const int ms = (lane >> 2) & 1, mk = lane & 3;
        float cv = INFINITY; int cb = 0x7fffffff, co = 0;
        if (lane < 8) { cv = res->cval[ms][mk][i]; cb = res->cbin[ms][mk][i]; co = res->cord[ms][mk][i]; }
        int rank = 0;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
          const float ovv = __shfl_sync(FULL, cv, l); const int obb = __shfl_sync(FULL, cb, l);
          rank += (ovv < cv || (ovv == cv && obb < cb)) ? 1 : 0;
        }
        Merged m; m.val = 0.f; m.bin = 0; m.best_ord = 0;
        m.nvalid = res->nem[0][i] + res->nem[1][i];
        int best_tile_ord = 0;
#pragma unroll
        for (int r = 0; r < ST_KL; ++r) {
          const unsigned who = __ballot_sync(FULL, lane < 8 && rank == r);
          const int src = who ? (__ffs(who) - 1) : 0;
          const float sv = __shfl_sync(FULL, cv, src); const int sb = __shfl_sync(FULL, cb, src); const int so = __shfl_sync(FULL, co, src);
          if (lane == r) { m.val = sv; m.bin = sb; }
          if (r == 0) { if (lane == 0) m.val = sv; best_tile_ord = so; }
        }
        if (m.nvalid > 0 && m.nvalid < K) {
          // the reference's fill-in uses the POSITION of the best peak in the peak list: peaks in earlier bin tiles + its ordinal
          const int bb = __shfl_sync(FULL, m.bin, 0);
          const int tb = bb / ST_STRIDE;
          int acc = 0;
          for (int t = lane; t < tb; t += 32) acc += cnt[t * ST_FRAMES + i];
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
          m.best_ord = acc + best_tile_ord - 1;
        }
        auto q_at = [&](int bin) -> float {   // only a frame without any local minimum evaluates the polynomial again (Horner form)
          const float2 dummy[1] = {make_float2(0.f, 0.f)};
          return q_coarse<0>(dummy, us, M, zplain[bin]);
        };
        if (m.nvalid == 0) { __syncwarp(); for (int l = lane; l < M; l += 32) us[l] = u[f * M + l]; __syncwarp(); }
        peaks_refine_emit<MT, true>(m, res->q0[i], res->qe[i], q_at, DirectEval<MT, true>{Gf, Vtab, M}, xaxis, M, P, K, lane, ov, ol, ob);
      }
      __syncwarp();
      if (more) g_park(buf ^ 1);
    }
    __syncthreads();                          // res is rewritten by the next tile
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == ST_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(512));
}

}  // namespace

// Host side: the steering-power table B[i][k] = (1, cos psi_i, sin psi_i, cos 2 psi_i, ...) for the plan's grid, split into
// tf32 hi / lo parts from float64 and laid out as the shared-memory images of its bin tiles (tile t, row r = bin 125 t - 1 + r,
// the last tile starting at bin P - 126; clamped to the grid: the copies at both ends can never look like peaks):
//   tile t: [hi image 16 KB][lo image 16 KB];  image byte of (row r, k) = (r >> 3) * 1024 + (r & 7) * 128 + (((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4.
// psi_i is the per-element phase step of the reference's steering table (build_music_tables): psi = -s d with the float s
// the constructor computes (lib/MUSIC_lin_array_impl.cc:103).
int scan_tc_ksteps(int M) { return (2 * M - 1 + 7) / 8; }
bool scan_tc_covers(int M, int P, int K) { return M >= 2 && M <= 16 && P >= 126 && P <= 125 * 255 && K >= 1 && K <= ST_KL; }

void build_scan_tc_table(float norm_spacing, int M, int P, const std::vector<float>& theta, std::vector<float>& out) {
  const double pi = 3.14159265358979323846;
  const int nbt = (P + ST_STRIDE - 1) / ST_STRIDE;
  out.assign((size_t)nbt * ST_STAGE_BYTES / sizeof(float), 0.0f);
  auto tf32 = [](float x) {
    uint32_t b; std::memcpy(&b, &x, 4);
    b = (b + 0x1000u) & 0xFFFFE000u;
    float r; std::memcpy(&r, &b, 4);
    return r;
  };
  for (int t = 0; t < nbt; ++t) {
    float* hi_img = out.data() + (size_t)t * (ST_STAGE_BYTES / 4);
    float* lo_img = hi_img + ST_TILE_BYTES / 4;
    for (int r = 0; r < ST_BINS; ++r) {
      const int i = std::min(std::max((t == nbt - 1 ? P - 126 : t * ST_STRIDE - 1) + r, 0), P - 1);
      const float s = (float)(-1.0 * 2 * pi * std::cos((double)theta[i]));
      const double psi = -(double)s * (double)norm_spacing;
      for (int k = 0; k < 2 * M - 1; ++k) {
        const int l = (k + 1) / 2;
        const double v = (k == 0) ? 1.0 : ((k & 1) ? std::cos(l * psi) : std::sin(l * psi));
        const float hi = tf32((float)v);
        const float lo = tf32((float)(v - (double)hi));
        const size_t off = ((size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 + ((((size_t)k >> 2) ^ (size_t)(r & 7)) << 4) + (size_t)(k & 3) * 4) / 4;
        hi_img[off] = hi;
        lo_img[off] = lo;
      }
    }
  }
}

namespace {
template <int MT>
int launch_tc_mt(const float2* u, const float2* G, const ScanTables& tb, int nframes, int K, float* out_val, float* out_loc, int* out_bin,
                 cudaStream_t st) {
  const int M = tb.M, nbt = (tb.P + ST_STRIDE - 1) / ST_STRIDE;
  const size_t smem = (size_t)ST_STAGES * ST_STAGE_BYTES + sizeof(TileResults) + (size_t)ST_WARPS * (2 * M * M + M) * sizeof(float2) +
                      (size_t)34 * 8 * 32 * sizeof(float) + (size_t)nbt * ST_FRAMES + 1024 + 16;
  if (smem > 225 * 1024) return 0;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = (nframes + ST_FRAMES - 1) / ST_FRAMES;
  const int grid = std::min(ntiles, sms);
  const int ks = scan_tc_ksteps(M), dbg = dev_option(OPT_SCAN_TC_DBG, 0);
  if (K == 1) {
    auto kern = scan_tc_kernel<MT, true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, ST_THREADS, smem, st>>>(u, G, tb.tctab, ks, tb.z, tb.zpair, tb.V, tb.xaxis, M, tb.P, nframes, K, out_val, out_loc, out_bin, dbg);
  } else {
    auto kern = scan_tc_kernel<MT, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, ST_THREADS, smem, st>>>(u, G, tb.tctab, ks, tb.z, tb.zpair, tb.V, tb.xaxis, M, tb.P, nframes, K, out_val, out_loc, out_bin, dbg);
  }
  return 1;
}
}  // namespace

int launch_scan_peaks_tc(const float2* u, const float2* G, const ScanTables& tb, int nframes, int K, float* out_val,
                         float* out_loc, int* out_bin, cudaStream_t st) {
  if (nframes <= 0) return 0;
  if (tb.tctab == nullptr || !scan_tc_covers(tb.M, tb.P, K)) return 0;
  if ((reinterpret_cast<uintptr_t>(G) & 15u) != 0 || (tb.M & 1)) return 0;   // the refinement stages G with 16-byte loads
  switch (tb.M) {
    case 4: return launch_tc_mt<4>(u, G, tb, nframes, K, out_val, out_loc, out_bin, st);
    case 8: return launch_tc_mt<8>(u, G, tb, nframes, K, out_val, out_loc, out_bin, st);
    case 16: return launch_tc_mt<16>(u, G, tb, nframes, K, out_val, out_loc, out_bin, st);
    default: return launch_tc_mt<0>(u, G, tb, nframes, K, out_val, out_loc, out_bin, st);
  }
}

}  // namespace doa
