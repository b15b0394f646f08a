// tcgen05 flash attention for the SigLIP vision tower (S = 729 tokens, 16 heads, dh = 72, non-causal).
//
// Reference op: softmax(Q K^T / sqrt(dh)) V per (image, head) inside each of the 27 blocks
// (aitemplate/model.py:30-36: nn.MultiheadAttention(use_mem_eff=True); clip_server.py:114).
//
// Persistent kernel, one CTA per SM, work item = 256 query rows (two 128-row tiles A / B) of one (image, head):
//   warp 8      TMA producer.  qkv is viewed as a 3-D tensor [token][3*H head slots][72]; a box of 64 dims lands as a
//               SWIZZLE_128B K-major tile and the remaining 8 dims as a 16-wide SWIZZLE_32B tile whose upper 8 columns
//               are out of bounds in dim 0 and therefore zero-filled by TMA -- dh = 72 is padded to 80 for free.  Four K/V stages.
//   warps 9/10  MMA issuers, one per query tile.  S_j = Q K_j^T: 4 x (K=16, SW128) + 1 x (K=16, SW32) tcgen05.mma, M=128,
//               N=128 keys, fp32 in TMEM.  O += P_j V_j: per 16 keys ONE N=80 MMA whose A operand P is read from TENSOR MEMORY
//               (TS form) and whose B operand V is five SW32 MN-major tiles exactly as V lies in memory ([key][dh]).
//   warps 0-3 / 4-7  softmax groups of tile A / B: thread = query row.  Reads S_j from TMEM once (128 scores in registers), exp2,
//               writes P_j as packed fp16 pairs with tcgen05.st into ONE 64-column P tile the two query tiles share, keeps the
//               running max / sum; O accumulates in TMEM with a lazy rescale.  Normalises and stores fp16 at the end.
// The two softmax groups alternate on a token (named barriers): one tile's exponentials (MUFU, 16/clk/SM: ~1.1 k cycles per
// 128 x 128 block) overlap the other tile's score load, P store and barrier round trips.  r03: P moved from shared memory (two SW128
// tiles per query tile, 128 B/clk + a generic-to-async proxy fence per block) to tensor memory (tcgen05.st, 256 B/clk, no proxy
// fence); a tile writes the shared P buffer only after the other tile's P V has read it, which the token order makes a wait that is
// already satisfied.  0.306-0.33 -> 0.277 ms per layer at batch 64 (565 TFLOP/s); 45.4 -> 40.4 ms inside the power-capped tower step.
// Measured without gain on top of it (profiles/r03f_attention_p_in_tmem.md): 2^x for a quarter / half of the elements on the FMA pipe
// (Cody-Waite polynomial; packed fp32 inside the token phase, and scalar OUTSIDE it -- every variant is slower the more elements it
// moves), and passing the token a quarter / half / three quarters of the exponentials early (no change at all).
#pragma once
#include "ptx.cuh"
#include <cuda_fp16.h>

namespace mse {
namespace attn_tc {

static constexpr int kBM = 128;          // query rows per tile; a work item is TWO tiles (256 rows) sharing every K/V block
static constexpr int kBN = 128;          // keys per block
static constexpr int kDH = 72;
static constexpr int kThreads = 384;     // warps 0-3: softmax group A, 4-7: softmax group B, 8: TMA, 9/10: MMA issuers (tile A/B), 11: idle
static constexpr int kKVStages = 4;       // the P tiles moved to tensor memory: their 64 KB hold a fourth K/V stage
static constexpr uint32_t kT64 = kBM * 128;   // [128 rows][64 halfs] SW128 tile bytes
static constexpr uint32_t kT16 = kBM * 32;    // [128 rows][16 halfs] SW32 tile bytes
static constexpr uint32_t kQBytes = 2 * (kT64 + kT16);      // both query tiles
static constexpr uint32_t kKVBytes = 2 * (kT64 + kT16);     // K tile pair + V tile pair
// smem map (all tile bases 1024-aligned)
static constexpr uint32_t kOffQ = 0;
static constexpr uint32_t kOffKV = kOffQ + kQBytes;                     // 40960
static constexpr uint32_t kOffBar = kOffKV + kKVStages * kKVBytes;      // + 163840 = 204800
static constexpr uint32_t kSmemBytes = kOffBar + 256 + 1024;
static constexpr uint32_t kTmemCols = 512;
// S_A @0, S_B @128 (fp32 scores), O_A @256, O_B @336 (80 fp32 columns each), P @416: ONE fp16 probability tile [128 rows][128 keys] = 64
// columns shared by the two query tiles -- they take turns (ping-pong), and a tile writes P only after the other tile's P V has read it
static constexpr uint32_t kTmS = 0, kTmO = 256, kTmOStride = 80, kTmP = 416;

// K-major SWIZZLE_32B tile (rows of 32 bytes, 8-row atoms of 256 bytes)
__device__ __forceinline__ uint64_t smem_desc_sw32(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// idesc with B taken MN-major (V as it lies: [key][dh], dh contiguous)
__host__ __device__ constexpr uint32_t idesc_f16_bmn(int m, int n) { return ptx::umma_idesc_f16(m, n, 0) | (1u << 16); }

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(ptx::smem_u32(smem_dst)), "l"((uint64_t)m), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M rows = lanes, K fp16 elements packed two per 32-bit column) is read from tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    // volatile: the exponentials must stay between the ping-pong barriers (pure arithmetic would be free to move across them)
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Params {
    int S, H, B;
    int q_items;        // ceil(S / 256) work items per (image, head)
    int n_blocks;       // ceil(S / 128) key blocks
    int n_items;        // B * H * q_items
    float scale_log2e;
    int debug;          // profiling only: 1 skip ex2, 2 skip P stores, 4 skip PV MMAs, 8 skip S MMAs, 16 skip K/V TMA loads, 64 no ping-pong
    long long *trace;   // profiling only: CTA 0 records clock64() at phase boundaries, [5 roles][256] (mse_debug_attention mode bit 256)
};

// tm64: box {64, 1, 128} SWIZZLE_128B; tm16: box {16, 1, 128} SWIZZLE_32B; both over qkv viewed as [B*S][3H][72]
__global__ void __launch_bounds__(kThreads, 1)
k_mha_tc(const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm16, __half *__restrict__ out, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(smem + kOffBar);
    uint64_t *q_full = bars + 0, *q_empty = bars + 1;
    uint64_t *kv_full = bars + 2, *kv_empty = bars + 2 + kKVStages;           // [kKVStages] each
    uint64_t *s_full = bars + 2 + 2 * kKVStages, *s_empty = s_full + 2;       // [2] = per query tile
    uint64_t *p_full = s_empty + 2, *o_full = p_full + 2, *o_empty = o_full + 2;  // [2] = per query tile
    uint32_t *tmem_slot = (uint32_t *)(o_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 8 && lane == 0) {
        ptx::prefetch_tensormap(&tm64);
        ptx::prefetch_tensormap(&tm16);
        ptx::mbar_init(q_full, 1);
        ptx::mbar_init(q_empty, 2);  // one commit per MMA issuer
        for (int i = 0; i < kKVStages; i++) { ptx::mbar_init(&kv_full[i], 1); ptx::mbar_init(&kv_empty[i], 2); }
        for (int i = 0; i < 2; i++) {
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&s_empty[i], 128);
            ptx::mbar_init(&p_full[i], 128);
            ptx::mbar_init(&o_full[i], 1);
            ptx::mbar_init(&o_empty[i], 128);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 9) {
        ptx::tmem_alloc(tmem_slot, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nb = p.n_blocks;
    int tr_n = 0;
    auto TR = [&](int role) {
        if (p.trace && blockIdx.x == 0 && tr_n < 256) p.trace[role * 256 + tr_n++] = clock64();
    };

    // register budget: the two softmax groups need their 128-score row in registers; the TMA / MMA group needs almost none
    if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 8) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t item_n = 0, kv_n = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, item_n++) {
                const int qi = item % p.q_items, bh = item / p.q_items, h = bh % p.H, b = bh / p.H;
                const int tok0 = b * p.S, q0 = tok0 + qi * 2 * kBM;
                ptx::mbar_wait(q_empty, (item_n & 1) ^ 1);
                ptx::mbar_expect_tx(q_full, kQBytes);
                tma_load_3d(smem + kOffQ, &tm64, q_full, 0, h, q0);
                tma_load_3d(smem + kOffQ + kT64, &tm16, q_full, 64, h, q0);
                tma_load_3d(smem + kOffQ + kT64 + kT16, &tm64, q_full, 0, h, q0 + kBM);
                tma_load_3d(smem + kOffQ + 2 * kT64 + kT16, &tm16, q_full, 64, h, q0 + kBM);
                for (int j = 0; j < nb; j++, kv_n++) {
                    const uint32_t st = kv_n % kKVStages, ph = (kv_n / kKVStages) & 1;
                    ptx::mbar_wait(&kv_empty[st], ph ^ 1);
                    if (p.debug & 16) { ptx::mbar_arrive(&kv_full[st]); continue; }
                    ptx::mbar_expect_tx(&kv_full[st], kKVBytes);
                    uint8_t *kv = smem + kOffKV + st * kKVBytes;
                    const int key0 = tok0 + j * kBN;
                    tma_load_3d(kv, &tm64, &kv_full[st], 0, p.H + h, key0);
                    tma_load_3d(kv + kT64, &tm16, &kv_full[st], 64, p.H + h, key0);
                    // V as five [128 keys][16 dims] SWIZZLE_32B tiles, 4 KB apart: one MN-major B operand of N = 80 for the P V MMA
#pragma unroll
                    for (int c = 0; c < 5; c++) tma_load_3d(kv + kT64 + kT16 + c * kT16, &tm16, &kv_full[st], 16 * c, 2 * p.H + h, key0);
                }
            }
        }
    } else if (warp == 9 || warp == 10) {
        // ------------------------------------------------------------ MMA issuers: warp 9 drives query tile A, warp 10 tile B.
        // The MMAs of this kernel are small (N = 128 / 64 / 16), so the instruction stream of the issuing thread -- not the
        // tensor pipe -- sets the pace: two issuers run in parallel and every descriptor is a precomputed base plus a constant.
        if (lane == 0) {
            const uint32_t t = warp - 9;
            constexpr uint32_t idesc_s = ptx::umma_idesc_f16(kBM, kBN, 0);
            constexpr uint32_t idesc_o80 = idesc_f16_bmn(kBM, 80);
            // descriptor halves: lo = (addr >> 4) | LBO(1) << 16; hi = SBO >> 4 | version 1 << 14 | layout << 29
            constexpr uint32_t kHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);
            constexpr uint32_t kHi32 = (256u >> 4) | (1u << 14) | (6u << 29);
            auto lo_of = [&](uint32_t off) { return ((ptx::smem_u32(smem + off) & 0x3FFFFu) >> 4) | (1u << 16); };
            auto d64 = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
            const uint32_t q_lo = lo_of(kOffQ + t * (kT64 + kT16)), q2_lo = q_lo + (kT64 >> 4);
            const uint32_t kv_lo0 = lo_of(kOffKV);
            const uint32_t d_s = tmem_base + kTmS + t * kBN, d_o = tmem_base + kTmO + t * kTmOStride;
            uint32_t item_n = 0, kv_n = 0, blk_n = 0;  // blk_n: running count of key blocks (phase of the per-tile barriers)
            // S_t(block) = Q_t K^T into the tile's score buffer
            auto issue_s = [&](uint32_t kvn, uint32_t bn) {
                const uint32_t st = kvn % kKVStages, ph = (kvn / kKVStages) & 1;
                ptx::mbar_wait(&kv_full[st], ph);
                TR(2 + (int)t);                  // 0: K/V landed
                ptx::mbar_wait(&s_empty[t], (bn & 1) ^ 1);
                TR(2 + (int)t);                  // 1: S buffer free
                ptx::tc_fence_after();
                const uint32_t k_lo = kv_lo0 + st * (kKVBytes >> 4);
                if (!(p.debug & 8)) {
#pragma unroll
                for (uint32_t k = 0; k < 4; k++) ptx::umma_f16(d_s, d64(q_lo + 2 * k, kHi128), d64(k_lo + 2 * k, kHi128), idesc_s, k != 0);
                ptx::umma_f16(d_s, d64(q2_lo, kHi32), d64(k_lo + (kT64 >> 4), kHi32), idesc_s, 1u);
                }
                ptx::umma_commit(&s_full[t]);
                TR(2 + (int)t);                  // 2: S issued
            };
            // O_t += P_t(block) V(block): accumulates in TMEM across the key blocks of an item
            auto issue_pv = [&](uint32_t kvn, uint32_t bn, bool first_block, uint32_t itn) {
                const uint32_t st = kvn % kKVStages;
                ptx::mbar_wait(&p_full[t], bn & 1);
                TR(2 + (int)t);                  // 3: P ready
                if (first_block) ptx::mbar_wait(&o_empty[t], (itn & 1) ^ 1);  // previous item's output has been read out
                ptx::tc_fence_after();
                // V: five SW32 MN-major tiles [key][16 dh]; atoms of 8 keys x 32 B = 256 B follow each other along the keys (SBO = 256),
                // the next 16 dims are the next tile (LBO = 4096): ONE N = 80 MMA per 16 keys reads P once (the r01 kernel issued an
                // N = 64 and an N = 16 MMA, each fetching the whole P operand; the issuing thread, ~100 cycles per MMA, set the pace)
                const uint32_t v_lo = (kv_lo0 & 0xFFFFu) + st * (kKVBytes >> 4) + ((kT64 + kT16) >> 4);
                if (!(p.debug & 4))
#pragma unroll
                for (uint32_t ks = 0; ks < kBN / 16; ks++) {
                    const uint32_t accum = (!first_block || ks != 0) ? 1u : 0u;
                    // P from tensor memory: 16 keys = 8 columns of packed fp16 pairs
                    umma_f16_ts(d_o, tmem_base + kTmP + ks * 8, d64((v_lo + ks * (512 >> 4)) | ((kT16 >> 4) << 16), kHi32), idesc_o80, accum);   // +16 keys = 512 B
                }
                ptx::umma_commit(&o_full[t]);
                TR(2 + (int)t);                  // 4: PV issued
            };
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, item_n++) {
                ptx::mbar_wait(q_full, item_n & 1);
                issue_s(kv_n, blk_n);
                if (nb == 1) ptx::umma_commit(q_empty);
                for (int j = 0; j < nb; j++) {
                    if (j + 1 < nb) {
                        issue_s(kv_n + j + 1, blk_n + j + 1);
                        if (j + 2 == nb) ptx::umma_commit(q_empty);  // this tile's last QK^T is issued: Q may be reloaded once both tiles say so
                    }
                    issue_pv(kv_n + j, blk_n + j, j == 0, item_n);
                    ptx::umma_commit(&kv_empty[(kv_n + j) % kKVStages]);  // second of the two arrivals frees the stage
                }
                kv_n += nb;
                blk_n += nb;
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ------------------------------------------------------------ softmax / output groups (thread = query row)
        // TMEM -> register bandwidth (64 B/clk/SM) is the scarce resource, so every score is read exactly once: the thread
        // keeps its whole 128-key row in registers.  The output accumulates in TMEM across key blocks; it is rescaled in
        // place only when the running max has grown by more than 2^8 since the reference the P tiles are expressed in
        // (P <= 256 fits fp16 comfortably), which after the first block or two almost never happens.
        const uint32_t t = warp >> 2;  // query tile owned by this group
        const uint32_t quad = warp & 3;
        const uint32_t row = quad * 32 + lane;
        const uint32_t lane_base = (quad * 32) << 16;
        const float sc = p.scale_log2e;
        const uint32_t ts = tmem_base + lane_base + kTmS + t * kBN;
        const uint32_t to = tmem_base + lane_base + kTmO + t * kTmOStride;
        const uint32_t tp = tmem_base + lane_base + kTmP;
        uint32_t blk_n = 0, item_n = 0;
        // The exponentials are MUFU-bound (16 ex2/clk/SM: a tile's 128 x 128 block is 1024 cycles when it has the unit to itself) and
        // the rest of a block -- score load, P stores, barrier round trips -- is not.  Left alone the two tiles run in lockstep and
        // share the MUFU in the same phase (2 x ~2500 cycles per block, phase timeline in profiles/r02_attention_timeline.md); a token
        // passed between the two softmax groups on a pair of named barriers makes them alternate, so one tile's exponentials overlap
        // the other tile's loads and stores.
        const bool pingpong = true;   // the shared P buffer relies on the alternation
        if (pingpong && t == 1) ptx::named_bar_arrive(2, 256);   // tile A goes first
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, item_n++) {
            const int qi = item % p.q_items, bh = item / p.q_items, h = bh % p.H, b = bh / p.H;
            float m_ref = -INFINITY;  // reference max the P tiles / O / l are expressed against
            float l_run = 0.f;
            for (int j = 0; j < nb; j++) {
                const uint32_t bn = blk_n + j;
                ptx::mbar_wait(&s_full[t], bn & 1);
                ptx::tc_fence_after();
                if (row == 0) TR((int)t);        // 0: S ready
                uint32_t v[kBN];
#pragma unroll
                for (int c = 0; c < kBN / 32; c++) ptx::tmem_ld_32x32(ts + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&s_empty[t]);  // the score buffer can take block j+1 while we work from registers
                if (row == 0) TR((int)t);        // 1: scores in registers
                const int kvalid = p.S - j * kBN;   // keys of this block that exist (>= 1); only the last block is partial
                if (kvalid < kBN) {                 // warp-uniform
#pragma unroll
                    for (int i = 0; i < kBN; i++) v[i] = (i < kvalid) ? v[i] : 0xff800000u;  // -inf
                }
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                for (int i = 0; i < kBN; i += 4) {
                    mx0 = fmaxf(mx0, __uint_as_float(v[i]));
                    mx1 = fmaxf(mx1, __uint_as_float(v[i + 1]));
                    mx2 = fmaxf(mx2, __uint_as_float(v[i + 2]));
                    mx3 = fmaxf(mx3, __uint_as_float(v[i + 3]));
                }
                const float m_blk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;
                const bool grow = m_blk > m_ref + 8.0f;  // also true on the first block (m_ref = -inf)
                const float f = grow ? ex2_approx(m_ref - m_blk) : 1.0f;  // rescale of O and l if the reference moves
                if (grow) m_ref = m_blk;
                const float neg_m = -m_ref;
                // the row maximum above and the scaling of the scores need no MUFU: only the exponentials run under the token
#pragma unroll
                for (int i = 0; i < kBN; i++) v[i] = __float_as_uint(fmaf(__uint_as_float(v[i]), sc, neg_m));
                if (pingpong) ptx::named_bar_sync(2 + t, 256);     // wait for the MUFU token
                // all 128 exponentials first, packed to fp16 in registers: none of this needs the P buffer or O, so it
                // overlaps the P V MMAs of the previous block
                uint32_t pk[kBN / 2];
                float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
                for (int i = 0; i < kBN / 2; i++) {
                    float p0 = __uint_as_float(v[2 * i]), p1 = __uint_as_float(v[2 * i + 1]);
                    if (!(p.debug & 1)) { p0 = ex2_approx(p0); p1 = ex2_approx(p1); }
                    rs0 += p0;
                    rs1 += p1;
                    __half2 hh = __floats2half2_rn(p0, p1);
                    pk[i] = *(uint32_t *)&hh;
                }
                if (pingpong) ptx::named_bar_arrive(2 + (t ^ 1), 256);   // hand the token to the other tile
                if (row == 0) TR((int)t);        // 2: max + exps done
                if (j > 0) {
                    // P V of block j-1 must be complete before O is touched or the P buffer is overwritten
                    ptx::mbar_wait(&o_full[t], (bn - 1) & 1);
                    ptx::tc_fence_after();
                    if (__any_sync(0xffffffffu, grow)) {
                        uint32_t a[32], c2[32], d2[8];
                        ptx::tmem_ld_32x32(to, a);
                        ptx::tmem_ld_32x32(to + 32, c2);
                        tmem_ld_32x32_x8(to + 64, d2);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i++) a[i] = __float_as_uint(__uint_as_float(a[i]) * f);
#pragma unroll
                        for (int i = 0; i < 32; i++) c2[i] = __float_as_uint(__uint_as_float(c2[i]) * f);
#pragma unroll
                        for (int i = 0; i < 8; i++) d2[i] = __float_as_uint(__uint_as_float(d2[i]) * f);
                        tmem_st_32x32(to, a);
                        tmem_st_32x32(to + 32, c2);
                        tmem_st_32x32_x8(to + 64, d2);
                        tmem_st_wait();
                    }
                }
                if (row == 0) TR((int)t);        // 3: PV(j-1) done (+ rescale)
                l_run = fmaf(l_run, f, rs0 + rs1);
                // the P buffer is shared: the other tile's P V of its latest block must have read it.  The token order is A(bn), B(bn),
                // A(bn + 1), ...: tile B waits for P V_A(bn), tile A for P V_B(bn - 1) -- both long done in steady state.
                if (t == 1) ptx::mbar_wait(&o_full[0], bn & 1);
                else if (bn > 0) ptx::mbar_wait(&o_full[1], (bn - 1) & 1);
                ptx::tc_fence_after();
                // P as packed fp16 pairs straight into tensor memory (256 B/clk, against 128 B/clk + a proxy fence through shared memory):
                // thread = row = lane, 64 columns = 128 keys; the P V MMA takes its A operand from there
                if (!(p.debug & 2)) {
                    tmem_st_32x32(tp, *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]));
                    tmem_st_32x32(tp + 32, *reinterpret_cast<const uint32_t(*)[32]>(&pk[32]));
                    tmem_st_wait();
                }
                ptx::tc_fence_before();
                ptx::mbar_arrive(&p_full[t]);
                if (row == 0) TR((int)t);        // 4: P stored
            }
            // output: O / l for this row -> out[b*S + q][h*72 .. +72)
            ptx::mbar_wait(&o_full[t], (blk_n + nb - 1) & 1);
            ptx::tc_fence_after();
            {
                uint32_t a[32], c2[32], d2[8];
                ptx::tmem_ld_32x32(to, a);
                ptx::tmem_ld_32x32(to + 32, c2);
                tmem_ld_32x32_x8(to + 64, d2);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(&o_empty[t]);
                const int qrow = (qi * 2 + (int)t) * kBM + (int)row;
                if (qrow < p.S) {
                    const float inv = 1.f / l_run;
                    __half *dst = out + ((size_t)(b * p.S + qrow) * p.H + h) * kDH;
#pragma unroll
                    for (int c = 0; c < 9; c++) {
                        const uint32_t *src = c < 4 ? &a[c * 8] : (c < 8 ? &c2[(c - 4) * 8] : &d2[0]);
                        uint4 u;
                        __half2 *hh = (__half2 *)&u;
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            hh[e] = __floats2half2_rn(__uint_as_float(src[2 * e]) * inv, __uint_as_float(src[2 * e + 1]) * inv);
                        *(uint4 *)(dst + c * 8) = u;
                    }
                }
            }
            blk_n += nb;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace attn_tc
}  // namespace mse
