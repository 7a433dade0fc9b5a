// tcgen05 / TMEM tensor-core path of the fused emulator kernel (sm_100a).
//
// One persistent CTA per SM (a cluster of two works on a 256-row super-tile with cta_group::2 MMAs) processes 128-row tiles.  All
// Dense layers of a tile are chained on-chip: every layer is a sequence of tcgen05.mma (M = 128 per CTA, fp32 accumulate in TMEM)
// in a split product  D += A_hi*W_hi + A_hi*W_lo + A_lo*W_hi  -- three kind::f16 passes (bf16 or fp16 hi/lo pairs), or one
// kind::f16 pass plus one kind::f8f6f4 pass carrying both corrections --, the epilogue warps read the accumulators back with
// tcgen05.ld, add bias, apply ReLU, split the result again and hand it to the next layer either through shared memory (UMMA "SS"
// operand) or -- converted IN PLACE over the accumulator columns -- through TMEM (UMMA "TS" operand).  Weights (1.5 MB of packed
// operand images, L2 resident) stream through a ring of shared-memory slots filled by 1-D bulk (TMA) copies.  The first layer's
// operand comes from the fused parameter transform; the last layer's epilogue applies the output transform (or the chi^2
// reduction) and writes the spectra through tensor-map (TMA) box stores.
//
// Warp roles (128 + 128 EPS threads): role 0 = bulk-copy producer; roles 1 and 3 = MMA issuers, interpreting the host-built issue
// table (build_iters) record by record under a baton that fixes the issue order (role 1 also allocates TMEM; in the follower CTA of
// a pair it forwards "my half of this ring slot has landed" to the leader); role 2 = prologue (parameter transform -> layer-0
// operand, one tile ahead); roles 4.. = epilogue (thread = tile row = TMEM lane).
//
// History of the control path (profiles/README.md): the round-1 kernel walked layers and chunks in the issuing warp; measured in
// round 2, that lone warp was busy EXECUTING control instructions (chunk set-up, dependent constant loads, barrier bookkeeping) for
// 76 % of the launch.  The schedule is therefore flattened on the host into one 16-byte record per ring slot, and the output mode is a
// template parameter (each instantiation carries only its own final-layer code).
//
// Operand images (validated on hardware by tools/umma_probe.cu):
//   un-swizzled K-major core-matrix layout [k/8][row][8 x 16-bit]: descriptor LBO = bytes between
//   k-groups, SBO = 128 B between 8-row groups, version 1, layout type 0;
//   A in TMEM: lane = row, one 32-bit column per k pair (even k in the low half).
//
// Reference semantics: VeryAccurateEmulator/emulator.py:401-403 (see fp32_kernel.cuh for the
// bit-faithful path; this one is held to 0.01 mK rms / 0.05 mK max).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <cuda_fp8.h>

#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace tck {

constexpr int MT = 128;
#ifndef VAE21_TC_EPI_PER_SUB
#define VAE21_TC_EPI_PER_SUB 4   // epilogue warps per TMEM sub-partition (2, 3 or 4): 4 measured best (1.95 vs 1.99 ms)
#endif
constexpr int EPS = VAE21_TC_EPI_PER_SUB;
constexpr int NEPI = 128 * EPS;          // epilogue threads
constexpr int NTHREADS = 128 + NEPI;     // roles: 0 producer, 1 MMA issuer A (follower CTA of a pair: ring forwarder), 2 prologue, 3 MMA issuer B, 4.. epilogue
                                         // (a 21st warp would cost 16 registers per thread: the register file is allocated per 4 warps)
constexpr int WARP_ISSUER_B = 3;  // role index of the second MMA-issuing warp
constexpr int MAXL = 8;
constexpr int MAXC = 32;
constexpr int NFULL = 4;          // ring of "accumulator chunk ready" barriers
#ifndef VAE21_TC_ABLATE
#define VAE21_TC_ABLATE 0  // profiling only (bit mask): 1 no MMA issue, 2 no epilogue work, 4 no weight copies, 8 no final-layer stores,
                           // 16 no hi/lo split arithmetic, 32 no hidden-layer operand writes, 128 no weight ring at all (MMAs read stale smem)
#endif
constexpr int DBG = VAE21_TC_ABLATE;
#ifndef VAE21_TC_EPI_SINGLE
#define VAE21_TC_EPI_SINGLE 0  // 1: a single instantiation of the epilogue's group body (smaller code, 16 extra moves per group)
#endif
#ifndef VAE21_TC_STORE_DBG
#define VAE21_TC_STORE_DBG 0  // profiling only (bit mask): 1 no box store issue, 2 no scalar stores around the boxes, 4 no wait for earlier boxes, 8 no barrier
#endif
constexpr int SDBG = VAE21_TC_STORE_DBG;
#ifndef VAE21_TC_TIMING
#define VAE21_TC_TIMING 0  // profiling only: per-CTA cycle counters of the MMA warp's waits (tools/tc_timing.py)
#endif
#if VAE21_TC_TIMING
__device__ long long g_tc_timing[160][16];
__device__ unsigned long long g_tc_rec[3][256];  // per issue-table record, summed over CTAs: [0] flagged waits, [1] ring wait, [2] visits  // [cta][0 total, 1 a0 wait, 2 accumulator-free wait, 3 ring wait, 4 operand wait, 5 issue blocks, 6 loop iterations]
#endif
constexpr int MAX_SLOTS = 16;
#ifndef VAE21_TC_CTRL_LAST
#define VAE21_TC_CTRL_LAST 1  // control warps on the highest hardware warp ids (see the kernel)
#endif
#ifndef VAE21_TC_KPS
#define VAE21_TC_KPS 2   // k-steps per ring slot of the CTA-pair kernel (1 or 2; 1 measured slower: 2.05 vs 1.90 ms)
#endif
#ifndef VAE21_TC_WAIT_HINT
#define VAE21_TC_WAIT_HINT 0  // ns the hardware may suspend a warp inside mbarrier.try_wait (0 = no hint)
#endif
constexpr int BAR_BYTES = 512;  // barrier block at off_bar (2 * MAX_SLOTS + NFULL + 2 + MAX_LCHUNK + 1 barriers + the TMEM base word)
constexpr int MAX_LCHUNK = 4;      // chunks per non-final layer (per-chunk operand-ready barriers)
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int KSTEP_BYTES = 8192; // one k-step (16 features) of a 128-row activation tile: hi 4 KB + lo 4 KB
constexpr int A_KG_BYTES = 2048;  // 128 rows x 16 B: distance between the two k-groups of a k-step

enum : int { A_SMEM_A0 = 0, A_SMEM_ACT = 1, A_TMEM = 2 };
enum : int { DST_SMEM = 0, DST_TMEM = 1, DST_FINAL = 2 };

struct Layer {
    int K;         // input width padded to 16
    int N;         // true output width
    int Npad;      // padded to 16
    int relu;
    int a_src;     // A_*
    int out_dst;   // DST_*
    int bias_off;  // float offset into the bias image / smem copy
    int first_chunk, nchunks;
    float inv_s8;  // FMT 2 only: 1 / S_l, the power-of-two scale the packed weights (hence the accumulators) carry
};

struct Chunk {
    int layer;
    int n0, ncols;   // column range of the layer this chunk accumulates
    int dcol;        // TMEM column of the accumulator
    int qbuf;        // 0/1: ring buffer index, -1: in place (stays in TMEM as next layer's operand)
    int nstages;     // weight stages (each 16 k wide, hi+lo) == K/16
    int kps2;        // pair kernel: k-steps per ring slot for this chunk (as many as fit the slot, at most 4)
    // copies of the owning layer's fields, so the device loops read ONE record per chunk (no dependent constant loads)
    int idx_in_layer, last_in_layer;
    int a_src, out_dst, relu;
    int bias_n0;     // bias_off + n0
    int src_first, src_count;   // chunks of the producing layer (first chunk of a layer > 0 only, else src_count = 0)
    float inv_s8;
    unsigned w_off;  // byte offset of the first stage in the weight image
    int st_w, st_map;  // final-layer chunks, OM_ROWS: width of the chunk's tensor-store boxes (0: none) and which tensor map has that box
};

struct Plan {
    int n_layers, n_chunks;
    int K0;      // true number of input parameters
    int n_out;   // true output width
    int slot_bytes, nslots;      // weight ring, one CTA per tile (cta_group::1)
    int slot_bytes2, nslots2;    // weight ring of the CTA-pair kernel (cta_group::2: each CTA holds half of every B tile)
    int smem_total2;
    int default_cg;              // kernel variant used unless VAE21_TC_CTA_GROUP overrides it
    int issuers;                 // MMA-issuing warps: 2 (default; they alternate ring slots under a token, so the issue order is fixed) or 1
    int bias_total;
    int bias_smem;               // floats of the bias image copied to shared memory (the final layer's biases live in s_s0 only)
    // issue tables (one 16-byte record per ring slot of a tile, see build_iters): word offsets into the bias image, [0] one CTA per
    // tile, [1] CTA pairs
    int iter_off[2], n_iter[2];
    int l0_iters[2];             // ring slots of layer 0 (arrival count of the "a0 may be overwritten" barrier)
    int off_iter;                // shared-memory copy of the issue table
    // output staging (OM_ROWS, see the epilogue): per buffer 16 arrays (TMEM sub-partition x row residue mod 4) of stage_sq bytes,
    // each a dense [8 super-rows][box width] tile; 1 or 2 buffers; the distinct box widths (one tensor map each)
    int stage_sq, stage_bufs, n_maps, map_w[4];
    // shared-memory carve-up (bytes from the 1024-aligned base)
    int off_act, off_stage, off_a0, off_ring, off_ring2, off_bias, off_s0, off_obs, off_isig, off_bar, smem_total;
    unsigned w_bytes;
    Layer L[MAXL];
    Chunk C[MAXC];
};

// ---------------------------------------------------------------------------------------------
// Host: planning and weight packing
// ---------------------------------------------------------------------------------------------
inline unsigned short f2bf16(float x) {
    __nv_bfloat16 b = __float2bfloat16_rn(x);
    unsigned short u;
    std::memcpy(&u, &b, 2);
    return u;
}
inline float bf162f(unsigned short u) {
    unsigned int w = static_cast<unsigned int>(u) << 16;
    float f;
    std::memcpy(&f, &w, 4);
    return f;
}
inline unsigned short f2h16(float x) {
    __half h = __float2half_rn(x);
    unsigned short u;
    std::memcpy(&u, &h, 2);
    return u;
}
inline float h162f(unsigned short u) {
    __half h;
    std::memcpy(&h, &u, 2);
    return __half2float(h);
}
inline unsigned char f2e4m3(float x) { return static_cast<unsigned char>(__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3)); }
// FMT 2 ("fp16e4m3"): a*w ~= a_hi*w_hi [kind::f16] + e4m3(a)*e4m3(w_lo*2^11)*2^-11 + e4m3(a_lo*2^11)*2^-11*e4m3(w)  [one kind::f8f6f4
// MMA, K = 32 = the 16 features' {hi8 | lo8} against {w_lo8 ; w_hi8}].  The 2^-11 cannot be carried by an e4m3 operand (range),
// so the WHOLE accumulator is kept at scale S_l: w_hi is packed as fp16(w S_l), w_lo8 = e4m3(w S_l - w_hi), w_hi8 = e4m3(w S_l / 2^11)
// and the epilogue multiplies by 1 / S_l.  S_l = 2^11 unless a layer's weights are large (w S_l must stay inside fp16).
constexpr float A_LO_SCALE = 2048.f;

// ---- issue table -------------------------------------------------------------------------------
// The MMA-issuing warps do not walk layers and chunks: the host flattens a tile's schedule into one 16-byte record per ring slot
// ("iteration": 1..4 k-steps of one accumulator chunk), in exactly the order the weight producer fills the ring, and the issuers
// interpret the records.  Everything an iteration needs is in its record, so the two issuing warps can take alternate records
// without tracking each other's state (round-2 profile: with per-chunk set-up code and running counters the lone issuing warp was
// busy executing control instructions for 76 % of the kernel).
//   w0: A operand cursor (bits 0-15: TMEM column of the hi pairs, or shared-memory offset / 16 from the carve-up base) |
//       accumulator TMEM column << 16
//   w1: instruction-descriptor bits of the chunk width ((ncols >> 3) << 17)
//   w2: ncols (bits 0-8) | k-steps in the slot << 9 (3 bits) | flags << 12 | chunk index in the tile << 20
//   w3: barrier waits before the issue: byte 0 = accumulator buffer (IT_Q_*), bytes 1-3 = operand-ready barriers (IT_W_*)
enum : uint32_t {
    IT_TS = 1u << 12,        // A operand in TMEM
    IT_FIRST = 1u << 13,     // first k-step of the chunk: the first MMA overwrites the accumulator
    IT_COMMIT1 = 1u << 14,   // after the MMAs: arrive once / twice on the chunk's "accumulator complete" barrier (bits 14-15 = count; the
    IT_COMMIT2 = 2u << 14,   //   barrier expects two arrivals: the last two slots of a chunk -- possibly issued by different warps)
    IT_A0FREE = 1u << 16,    // layer 0: also arrive on "a0 may be overwritten"
};
// byte 0 of w3: bits 0-1 = 1 + accumulator ring buffer (0: none); bit 2 = parity of this use within the tile (already inverted: the
// wait is for the PREVIOUS use's drain); bit 3 = uses per tile odd (parity then alternates with the tile count); bit 4 = first use in
// the tile (nothing to wait for in the CTA's first tile)
// bytes 1-3 of w3: bits 0-5 = barrier index in the barrier block (0: none); bit 6 = parity of this use within the tile; bit 7 = uses
// per tile odd
constexpr int BAR_IDX_Q_EMPTY = 2 * MAX_SLOTS + NFULL;          // + buffer
constexpr int BAR_IDX_ACT_READY = 2 * MAX_SLOTS + NFULL + 2;    // + chunk index in the producing layer
constexpr int BAR_IDX_A0_READY = 2 * MAX_SLOTS + NFULL + 2 + MAX_LCHUNK;
constexpr int BAR_IDX_A0_FREE = BAR_IDX_A0_READY + 1;
static_assert(BAR_IDX_A0_FREE < 64, "barrier indices must fit the 6-bit fields of the issue table");

inline bool build_iters(const Plan& P, bool pair, std::vector<uint32_t>& tab, int& l0_iters, std::string& why) {
    tab.clear();
    l0_iters = 0;
    int U_q[2] = {0, 0}, U_act[MAX_LCHUNK] = {0, 0, 0, 0};  // uses per tile
    for (int c = 0; c < P.n_chunks; ++c)
        if (P.C[c].qbuf >= 0) ++U_q[P.C[c].qbuf];
    for (int l = 1; l < P.n_layers; ++l)
        for (int j = 0; j < P.L[l - 1].nchunks; ++j) ++U_act[j];
    int u_q[2] = {0, 0}, u_act[MAX_LCHUNK] = {0, 0, 0, 0};
    for (int c = 0; c < P.n_chunks; ++c) {
        const Chunk& C = P.C[c];
        const int kps = pair ? C.kps2 : 1, nst = C.nstages;
        const int niter = (nst + kps - 1) / kps;
        const bool ts = (C.a_src == A_TMEM);
        int src = -1, src_end = -1, next_src_k = 0x7fffffff;  // chunks of the producing layer still to wait for
        if (C.layer > 0 && C.idx_in_layer == 0) {
            src = C.src_first;
            src_end = src + C.src_count;
            next_src_k = 0;
        }
        for (int s = 0, i = 0; s < nst; s += kps, ++i) {
            const int nk = std::min(kps, nst - s);
            const uint32_t a = ts ? static_cast<uint32_t>(16 * s)
                                  : static_cast<uint32_t>(((C.a_src == A_SMEM_A0 ? P.off_a0 : P.off_act) + s * KSTEP_BYTES) >> 4);
            if (a > 0xffffu) { why = "internal: A cursor does not fit the issue table"; return false; }
            uint32_t w2 = static_cast<uint32_t>(C.ncols) | (static_cast<uint32_t>(nk) << 9) | (static_cast<uint32_t>(c) << 20);
            if (ts) w2 |= IT_TS;
            if (s == 0) w2 |= IT_FIRST;
            if (i == niter - 1) w2 |= (niter == 1) ? IT_COMMIT2 : IT_COMMIT1;
            else if (i == niter - 2) w2 |= IT_COMMIT1;
            if (C.layer == 0) {
                w2 |= IT_A0FREE;
                ++l0_iters;
            }
            uint32_t w3 = 0;
            if (s == 0 && C.qbuf >= 0) {
                const int u = u_q[C.qbuf]++;
                w3 |= static_cast<uint32_t>(1 + C.qbuf) | (static_cast<uint32_t>((u ^ 1) & 1) << 2) |
                      (static_cast<uint32_t>(U_q[C.qbuf] & 1) << 3) | (static_cast<uint32_t>(u == 0) << 4);
            }
            int nw = 0;
            auto add_wait = [&](int bar_idx, int u, int U) {
                ++nw;
                if (nw <= 3) w3 |= (static_cast<uint32_t>(bar_idx) | (static_cast<uint32_t>(u & 1) << 6) | (static_cast<uint32_t>(U & 1) << 7)) << (8 * nw);
            };
            if (s == 0 && C.layer == 0 && C.idx_in_layer == 0) add_wait(BAR_IDX_A0_READY, 0, 1);
            while (s + nk - 1 >= next_src_k) {  // the k-steps of this slot reach into the next chunk of the producing layer
                const int j = src - C.src_first;
                add_wait(BAR_IDX_ACT_READY + j, u_act[j]++, U_act[j]);
                ++src;
                next_src_k = (src < src_end) ? (P.C[src].n0 >> 4) : 0x7fffffff;
            }
            if (nw > 3) { why = "more than three operand waits in one ring slot"; return false; }
            tab.push_back(a | (static_cast<uint32_t>(C.dcol) << 16));
            tab.push_back(static_cast<uint32_t>(C.ncols >> 3) << 17);
            tab.push_back(w2);
            tab.push_back(w3);
        }
    }
    return true;
}

// Build the schedule for a Dense stack.  Returns false (with `why`) when the stack does not fit.
inline bool build_plan_with(bool first_to_tmem, int n_layers, const int* dims, const float* const* kernels,
                            const float* const* biases, const int* relu, Plan& P,
                            std::vector<unsigned short>* img /*[3]: bf16, fp16, fp16+e4m3*/, std::vector<float>& bias_img,
                            std::string& why) {
    P = Plan{};
    if (n_layers < 2 || n_layers > MAXL) { why = "needs 2.." + std::to_string(MAXL) + " layers"; return false; }
    if (relu[n_layers - 1]) { why = "last layer must be linear"; return false; }
    if (dims[0] > 16) { why = "more than 16 input parameters"; return false; }
    P.n_layers = n_layers;
    P.default_cg = 2;
    P.K0 = dims[0];
    P.n_out = dims[n_layers];
    auto pad16 = [](int x) { return (x + 15) / 16 * 16; };

    int smem_w = 0, tmem_w = 0;  // widest activation resident in smem / tmem
    int boff = 0;
    for (int l = 0; l < n_layers; ++l) {
        Layer& L = P.L[l];
        L.K = pad16(dims[l]);
        L.N = dims[l + 1];
        L.Npad = pad16(L.N);
        L.relu = relu[l] ? 1 : 0;
        // Operand placement.  Activations alternate between shared memory (UMMA "SS") and TMEM (converted in
        // place over the accumulator, UMMA "TS").  A layer that reads shared memory and fits ONE accumulator
        // chunk may also write shared memory (its epilogue starts after all its MMAs have read the input),
        // which leaves all of TMEM to the next layer's accumulators: fewer, wider MMAs.  Measured on the
        // DirectEmulator stack this is SLOWER (2.72 vs 2.52 ms per 1M rows: the wide last chunk's epilogue
        // delays the next tile), so it is opt-in (VAE21_TC_SS_LAST=1).
        if (l == 0) {
            L.a_src = A_SMEM_A0;
        } else {
            L.a_src = (P.L[l - 1].out_dst == DST_SMEM) ? A_SMEM_ACT : A_TMEM;
        }
        if (l == n_layers - 1) {
            L.out_dst = DST_FINAL;
        } else if (L.a_src == A_SMEM_A0) {
            L.out_dst = first_to_tmem ? DST_TMEM : DST_SMEM;  // the parity of the whole chain (see build_plan)
        } else if (L.a_src == A_SMEM_ACT) {
            L.out_dst = (L.Npad <= 224 && l == n_layers - 2 && std::getenv("VAE21_TC_SS_LAST")) ? DST_SMEM : DST_TMEM;
        } else {
            L.out_dst = DST_SMEM;
        }
        L.bias_off = boff;
        boff += L.Npad;
        if (L.out_dst == DST_SMEM) smem_w = std::max(smem_w, L.Npad);
        if (L.out_dst == DST_TMEM) tmem_w = std::max(tmem_w, L.Npad);
        if (l > 0 && L.K != P.L[l - 1].Npad) { why = "internal: width mismatch"; return false; }
    }
    P.bias_total = boff;
    if (tmem_w > 480) { why = "a TMEM-resident hidden layer is wider than 480"; return false; }

    // accumulator placement
    const int last = n_layers - 1;
    auto qgeom = [&](int l, int& q0, int& qsize, int& nbuf) {
        // columns free for ring accumulators while layer l runs: everything above its TMEM operand.
        // Measured MMA cost (tools/umma_probe.cu): A from TMEM max(96, N/2 + 10) cycles, A from smem
        // N/2 + 43 -- narrow accumulators waste the tensor pipe, so when two buffers would be
        // narrower than 112 columns use a single wide one (the MMA then alternates with the epilogue).
        const int lo = (P.L[l].a_src == A_TMEM) ? P.L[l].K : 0;
        const int freec = 512 - lo;
        nbuf = 2;
        qsize = std::min(256, (freec / 2) / 16 * 16);
        static const int min_q = std::getenv("VAE21_TC_MINQ") ? std::atoi(std::getenv("VAE21_TC_MINQ")) : 112;
        if (qsize < min_q) {
            nbuf = 1;
            qsize = std::min(256, freec / 16 * 16);
        }
        q0 = 512 - nbuf * qsize;
    };
    int nchunks = 0;
    unsigned woff = 0;
    int qflip = 0;
    for (int l = 0; l < n_layers; ++l) {
        Layer& L = P.L[l];
        L.first_chunk = nchunks;
        const int units = L.Npad / 16;
        int maxcols, q0 = 0, qsize = 0, nbuf = 2;
        if (L.out_dst == DST_TMEM) {
            static const int inplace_max = std::getenv("VAE21_TC_INPLACE_MAX") ? std::atoi(std::getenv("VAE21_TC_INPLACE_MAX")) : 224;
            maxcols = std::max(16, std::min(224, inplace_max / 16 * 16));  // stage = ncols * 16 k * 2 B * (hi + lo) <= 14336 B
            if (L.Npad > 512) { why = "in-place layer too wide"; return false; }
        } else {
            // layer 0 shares the ring geometry of the last layer (their accumulators overlap in
            // time across consecutive tiles without a drain in between)
            qgeom(l == 0 ? last : l, q0, qsize, nbuf);
            maxcols = std::min(qsize, 240);
            if (maxcols < 16) { why = "no TMEM left for accumulators"; return false; }
        }
        const int nch = (L.Npad + maxcols - 1) / maxcols;
        int done_units = 0;
        for (int c = 0; c < nch; ++c) {
            if (nchunks >= MAXC) { why = "too many accumulator chunks"; return false; }
            Chunk& C = P.C[nchunks];
            const int u = (units - done_units + (nch - c) - 1) / (nch - c);  // balanced, larger first
            C.layer = l;
            C.n0 = done_units * 16;
            C.ncols = u * 16;
            if (L.out_dst == DST_TMEM) {
                C.qbuf = -1;
                C.dcol = C.n0;
            } else {
                const int b = (nbuf == 1) ? 0 : qflip;
                C.qbuf = b;
                C.dcol = q0 + b * qsize;
                if (nbuf == 2) qflip ^= 1;
            }
            C.nstages = L.K / 16;
            C.w_off = woff;
            woff += static_cast<unsigned>(C.nstages) * C.ncols * 16 * 2 * 2;
            done_units += u;
            ++nchunks;
        }
        L.nchunks = nch;
        if (l < n_layers - 1 && nch > MAX_LCHUNK) { why = "too many chunks in a hidden layer"; return false; }
    }
    // every tile must use each ring buffer an even... no: parity is tracked with running counters.
    P.n_chunks = nchunks;
    P.w_bytes = woff;
    P.slot_bytes = 0;
    for (int c = 0; c < nchunks; ++c) P.slot_bytes = std::max(P.slot_bytes, P.C[c].ncols * 16 * 2 * 2);

    // ring slot geometry (independent of the shared-memory carve-up).  Pair kernel: each CTA holds half of every B tile, so a slot
    // of the same size holds TWO k-steps: the same bytes in flight with half as many barrier round trips (the ring protocol is
    // latency-, not bandwidth-bound); narrow chunks would leave most of a slot empty (112 columns: half), so up to 4 k-steps of
    // this CTA's half tile are packed into a slot.
    P.slot_bytes2 = P.slot_bytes / 2 * VAE21_TC_KPS;
    static const int kps_max = std::getenv("VAE21_TC_KPS_MAX") ? std::min(4, std::atoi(std::getenv("VAE21_TC_KPS_MAX"))) : 4;  // the issue loop unrolls 4
    for (int c = 0; c < nchunks; ++c)
        P.C[c].kps2 = std::max(VAE21_TC_KPS, std::min(std::max(kps_max, VAE21_TC_KPS), P.slot_bytes2 / (P.C[c].ncols * 32)));
    for (int c = 0; c < nchunks; ++c) {  // denormalised per-chunk records (the issue tables below read them)
        Chunk& C = P.C[c];
        const Layer& L = P.L[C.layer];
        C.idx_in_layer = c - L.first_chunk;
        C.last_in_layer = (c == L.first_chunk + L.nchunks - 1) ? 1 : 0;
        C.a_src = L.a_src;
        C.out_dst = L.out_dst;
        C.relu = L.relu;
        C.bias_n0 = L.bias_off + C.n0;
        C.src_first = 0;
        C.src_count = 0;
        if (C.layer > 0 && C.idx_in_layer == 0) {
            C.src_first = P.L[C.layer - 1].first_chunk;
            C.src_count = P.L[C.layer - 1].nchunks;
        }
    }
    P.n_iter[0] = P.n_iter[1] = 0;
    for (int c = 0; c < nchunks; ++c) {
        P.n_iter[0] += P.C[c].nstages;
        P.n_iter[1] += (P.C[c].nstages + P.C[c].kps2 - 1) / P.C[c].kps2;
    }

    // shared memory
    int off = 0;
    P.off_act = off;
    // Output staging (see the epilogue).  Box width of a final chunk: a multiple of 4 words that leaves room for the per-row shift
    // (<= 3 columns) inside the chunk's valid columns, and = 12 (mod 16): with that pitch the 32 rows of a warp hit 32 different banks.
    P.n_maps = 0;
    int wmax = 0;
    for (int c = P.L[last].first_chunk; c < nchunks; ++c) {
        const int valid = std::min(P.C[c].ncols, P.n_out - P.C[c].n0);
        int w = (valid - 3 >= 12) ? ((valid - 3 - 12) / 16 * 16 + 12) : 0;
        int mi = -1;
        for (int i = 0; i < P.n_maps; ++i)
            if (P.map_w[i] == w) mi = i;
        if (w > 0 && mi < 0) {
            if (P.n_maps < 4) {
                mi = P.n_maps;
                P.map_w[P.n_maps++] = w;
            } else {
                w = 0;  // more than four distinct widths: this chunk is written with plain stores
            }
        }
        P.C[c].st_w = w;
        P.C[c].st_map = mi < 0 ? 0 : mi;
        wmax = std::max(wmax, w);
    }
    P.stage_sq = (8 * std::max(wmax, 12) * 4 + 127) / 128 * 128;
    const int stage_buf_bytes = 16 * P.stage_sq;
    // The staging buffers alias the activation buffer where the last layer does not read it: all of it when the last layer's operand
    // is in TMEM, else the part above that operand (grown if needed).  Two buffers (one final chunk staged while the previous one is
    // being read out) where the activation buffer is that large anyway, else one.
    const int act_need = smem_w / 16 * KSTEP_BYTES;
    const int used = (P.L[last].a_src == A_SMEM_ACT) ? P.L[last].K / 16 * KSTEP_BYTES : 0;
    P.stage_bufs = (used + 2 * stage_buf_bytes <= std::max(act_need, 96 * 1024)) ? 2 : 1;
    const int act_bytes = std::max(act_need, used + P.stage_bufs * stage_buf_bytes);
    P.off_stage = off + used;
    off += act_bytes;
    P.off_a0 = off;
    off += KSTEP_BYTES;
    P.off_bias = off;
    P.bias_smem = P.L[last].bias_off;  // hidden layers only: the final layer's biases are folded into s_s0
    off += P.bias_smem * 4;
    const int nop = pad16(P.n_out);
    P.off_s0 = off;
    off += nop * 4;
    P.off_obs = off;
    off += nop * 4;
    P.off_isig = off;
    off += nop * 4;
    P.off_bar = off;
    off += BAR_BYTES + 2 * 16 * 4 + 4 * (EPS - 1) * 128 * 4;  // barriers + prologue constants + chi^2 / amplitude partials
    P.off_iter = off;  // issue table of the kernel variant that runs, then its weight ring
    const int off1 = (off + P.n_iter[0] * 16 + 127) / 128 * 128, off2 = (off + P.n_iter[1] * 16 + 127) / 128 * 128;
    P.off_ring = off1;
    P.off_ring2 = off2;
    // Even slot counts: with two issuing warps alternating slots, every use of a slot must be observed by the SAME warp (mbarrier
    // waits are by phase parity: a warp that saw only every other use of a slot would pass on a stale completion).
    P.nslots = std::max(0, std::min(MAX_SLOTS, (SMEM_LIMIT - 128 /*alignment slack*/ - off1) / P.slot_bytes) & ~1);  // 0: no one-CTA variant
    P.nslots2 = std::max(0, std::min(MAX_SLOTS, (SMEM_LIMIT - 128 - off2) / P.slot_bytes2) & ~1);
    if (P.nslots2 < 2) { why = "shared memory: activations leave no room for a weight ring"; return false; }
    P.smem_total2 = off2 + P.nslots2 * P.slot_bytes2 + 128;
    P.smem_total = off1 + P.nslots * P.slot_bytes + 128;
    static const int issuers_env = std::getenv("VAE21_TC_ISSUERS") ? std::atoi(std::getenv("VAE21_TC_ISSUERS")) : 2;
    P.issuers = issuers_env == 1 ? 1 : 2;

    // Weight images, in exactly the order the MMA warp consumes them: chunk -> k-step -> {hi, lo} tile,
    // tile = [2 k-groups][rows][8 elements]  (B operand, "K-major": row n holds W[k][n]).
    // Image 1 (bytes [0, w_bytes)): whole tiles for the one-CTA kernel.
    // Image 2 (bytes [w_bytes, 2 w_bytes)): for the CTA-pair kernel, rank r's half (rows [r n/2, (r+1) n/2) of
    // every tile) at w_bytes + r * w_bytes / 2, same order.
    for (int f = 0; f < 3; ++f) img[f].assign(P.w_bytes, 0);
    unsigned char* img8 = reinterpret_cast<unsigned char*>(img[2].data());
    for (int l = 0; l < n_layers; ++l) {
        float wmax = 0.f;
        for (size_t i = 0; i < static_cast<size_t>(dims[l]) * dims[l + 1]; ++i) wmax = std::max(wmax, std::fabs(kernels[l][i]));
        float S = A_LO_SCALE;
        while (S > 1.f && wmax * S > 32768.f) S *= 0.5f;
        P.L[l].inv_s8 = 1.f / S;
    }
    for (int c = 0; c < nchunks; ++c) P.C[c].inv_s8 = P.L[P.C[c].layer].inv_s8;

    const size_t img2 = P.w_bytes / 2;  // element offset of image 2
    for (int c = 0; c < nchunks; ++c) {
        const Chunk& C = P.C[c];
        const int l = C.layer;
        const int Kt = dims[l], Nt = dims[l + 1];
        const int hn = C.ncols / 2;
        for (int s = 0; s < C.nstages; ++s) {
            const size_t base = (C.w_off + static_cast<size_t>(s) * C.ncols * 64) / 2;  // in elements
            const size_t lo_base = base + static_cast<size_t>(C.ncols) * 16;
            for (int kk = 0; kk < 16; ++kk) {
                const int k = s * 16 + kk;
                for (int nn = 0; nn < C.ncols; ++nn) {
                    const int n = C.n0 + nn;
                    const float w = (k < Kt && n < Nt) ? kernels[l][static_cast<size_t>(k) * Nt + n] : 0.f;
                    const size_t e = (static_cast<size_t>(kk >> 3) * C.ncols + nn) * 8 + (kk & 7);
                    const int r = nn / hn, nl = nn - r * hn;
                    const size_t base2 = img2 + static_cast<size_t>(r) * (P.w_bytes / 4) + C.w_off / 4 +
                                         static_cast<size_t>(s) * C.ncols * 16;
                    const size_t lo2 = base2 + static_cast<size_t>(hn) * 16;
                    const size_t e2 = (static_cast<size_t>(kk >> 3) * hn + nl) * 8 + (kk & 7);
                    const unsigned short hb = f2bf16(w), lb = f2bf16(w - bf162f(hb));
                    img[0][base + e] = hb;
                    img[0][lo_base + e] = lb;
                    img[0][base2 + e2] = hb;
                    img[0][lo2 + e2] = lb;
                    const unsigned short hh = f2h16(w), lh = f2h16(w - h162f(hh));
                    img[1][base + e] = hh;
                    img[1][lo_base + e] = lh;
                    img[1][base2 + e2] = hh;
                    img[1][lo2 + e2] = lh;
                    // fp16 + e4m3: 16-bit tile = fp16(w S); 8-bit tile = k-group 0: e4m3(w S - hi)[16 k], k-group 1: e4m3(w S / 2^11)[16 k]
                    const float S = 1.f / P.L[l].inv_s8, ws = w * S;
                    const unsigned short h8 = f2h16(ws);
                    const unsigned char wl8 = f2e4m3(ws - h162f(h8)), wh8 = f2e4m3(ws / A_LO_SCALE);
                    img[2][base + e] = h8;
                    img[2][base2 + e2] = h8;
                    img8[2 * lo_base + static_cast<size_t>(nn) * 16 + kk] = wl8;
                    img8[2 * lo_base + (static_cast<size_t>(C.ncols) + nn) * 16 + kk] = wh8;
                    img8[2 * lo2 + static_cast<size_t>(nl) * 16 + kk] = wl8;
                    img8[2 * lo2 + (static_cast<size_t>(hn) + nl) * 16 + kk] = wh8;
                }
            }
        }
    }
    bias_img.assign(P.bias_total, 0.f);
    for (int l = 0; l < n_layers; ++l)
        for (int n = 0; n < dims[l + 1]; ++n) bias_img[P.L[l].bias_off + n] = biases[l][n];
    // issue tables, appended to the bias image (raw 32-bit words)
    for (int v = 0; v < 2; ++v) {
        std::vector<uint32_t> tab;
        if (!build_iters(P, v == 1, tab, P.l0_iters[v], why)) return false;
        if (static_cast<int>(tab.size()) != 4 * P.n_iter[v]) { why = "internal: issue table size"; return false; }
        P.iter_off[v] = static_cast<int>(bias_img.size());
        bias_img.resize(bias_img.size() + tab.size());
        std::memcpy(bias_img.data() + P.iter_off[v], tab.data(), tab.size() * 4);
    }
    return true;
}

// Activations alternate shared memory / TMEM along the chain; which of the two the FIRST hidden layer uses
// decides where the wide layers land.  Try "h1 in shared memory" (best for the DirectEmulator stack), and
// fall back to "h1 in TMEM" (the AE chain: its 352-wide layers then fit, and its last layer reads TMEM so
// the output staging can alias the activation buffer).  Keep the variant with the deeper weight ring.
inline bool build_plan(int n_layers, const int* dims, const float* const* kernels, const float* const* biases,
                       const int* relu, Plan& P, std::vector<unsigned short>* img, std::vector<float>& bias_img,
                       std::string& why) {
    std::string why_a, why_b;
    if (build_plan_with(false, n_layers, dims, kernels, biases, relu, P, img, bias_img, why_a) && P.nslots2 >= 4) return true;
    Plan Pa = P;
    const bool ok_a = why_a.empty() && Pa.n_chunks > 0 && Pa.nslots2 >= 2;
    std::vector<unsigned short> img_b[3];
    std::vector<float> bias_b;
    Plan Pb;
    const bool ok_b = build_plan_with(true, n_layers, dims, kernels, biases, relu, Pb, img_b, bias_b, why_b);
    if (ok_b && (!ok_a || Pb.nslots2 > Pa.nslots2)) {
        P = Pb;
        for (int f = 0; f < 3; ++f) img[f].swap(img_b[f]);
        bias_img.swap(bias_b);
        return true;
    }
    if (ok_a) return build_plan_with(false, n_layers, dims, kernels, biases, relu, P, img, bias_img, why_a);
    why = why_a.empty() ? why_b : why_a;
    return false;
}

// ---- schedule self-check (host only) ------------------------------------------------------------
// Replays the issue table of a plan against the epilogue's arrivals for a few tiles and verifies what the device protocol relies on:
//   * the records of every chunk cover its k-steps exactly once, in order, the first one overwriting the accumulator;
//   * every chunk's "accumulator complete" barrier gets exactly two commits, on its last one or two records; layer 0 commits the
//     "a0 may be overwritten" barrier once per record and the barrier is initialised with that count;
//   * every flagged wait names the phase it means: mbarrier waits are by phase PARITY, so at the time of a wait the previous phase
//     of that barrier must be known complete (an earlier record waited for it) and the phase after the awaited one must not be able to
//     complete before this record is issued -- otherwise the wait passes on a stale phase or blocks for ever.
// Run by vae21_check_plan (C ABI, no GPU needed) in the CPU tests over many layer stacks.
inline bool check_schedule(const Plan& P, const std::vector<float>& bias_img, std::string& why) {
    for (int v = 0; v < 2; ++v) {
        if (v == 0 && P.nslots < 2) continue;  // the one-CTA variant is optional
        const int n_iter = P.n_iter[v];
        const uint32_t* tab = reinterpret_cast<const uint32_t*>(bias_img.data()) + P.iter_off[v];
        if (P.iter_off[v] + 4 * n_iter > static_cast<int>(bias_img.size())) { why = "issue table outside the bias image"; return false; }
        // --- per-chunk structure
        std::vector<int> first_rec(P.n_chunks, -1), last_rec(P.n_chunks, -1);
        int r = 0, l0 = 0;
        for (int c = 0; c < P.n_chunks; ++c) {
            const Chunk& C = P.C[c];
            const int kps = v ? C.kps2 : 1;
            int k = 0, commits = 0;
            first_rec[c] = r;
            while (k < C.nstages) {
                if (r >= n_iter) { why = "issue table too short"; return false; }
                const uint32_t w0 = tab[4 * r], w1 = tab[4 * r + 1], w2 = tab[4 * r + 2];
                const int nk = (w2 >> 9) & 7;
                if (static_cast<int>(w2 >> 20) != c || static_cast<int>(w2 & 0x1ff) != C.ncols || nk < 1 || nk > kps || nk > 4) { why = "record does not match its chunk"; return false; }
                if (((w2 & IT_FIRST) != 0) != (k == 0)) { why = "accumulate flag"; return false; }
                if (((w2 & IT_TS) != 0) != (C.a_src == A_TMEM)) { why = "operand source flag"; return false; }
                if (static_cast<int>(w0 >> 16) != C.dcol || w1 != (static_cast<uint32_t>(C.ncols >> 3) << 17)) { why = "accumulator column / descriptor bits"; return false; }
                const uint32_t a = w0 & 0xffffu;
                const uint32_t want_a = (C.a_src == A_TMEM) ? 16u * k : static_cast<uint32_t>(((C.a_src == A_SMEM_A0 ? P.off_a0 : P.off_act) + k * KSTEP_BYTES) >> 4);
                if (a != want_a) { why = "A operand cursor"; return false; }
                const int nc = (w2 >> 14) & 3;
                commits += nc;
                if (nc && k + nk < C.nstages && k + nk + kps < C.nstages) { why = "accumulator commit before the last two records"; return false; }
                if (((w2 & IT_A0FREE) != 0) != (C.layer == 0)) { why = "a0-free flag"; return false; }
                if (C.layer == 0) ++l0;
                k += nk;
                ++r;
            }
            last_rec[c] = r - 1;
            if (k != C.nstages || commits != 2) { why = "k-steps / commit count of a chunk"; return false; }
        }
        if (r != n_iter || l0 != P.l0_iters[v]) { why = "record count"; return false; }
        // --- barrier phases over three tiles.  Producer events per barrier, in the epilogue's (= chunk) order; event e of a barrier
        // can only have happened once the last record of its chunk has been issued.
        const int T = 3;
        struct Ev { long long after_rec; };  // global index of the record that must have been issued before the event can happen
        std::vector<std::vector<Ev>> ev(64);
        for (int t = 0; t < T; ++t) {
            // the prologue runs one tile ahead: a0 of tile t is written once the layer-0 MMAs of tile t - 1 have released it
            {
                int l0_last = 0;
                for (int c = 0; c < P.n_chunks; ++c)
                    if (P.C[c].layer == 0) l0_last = last_rec[c];
                ev[BAR_IDX_A0_READY].push_back({t == 0 ? -1 : static_cast<long long>(t - 1) * n_iter + l0_last});
            }
            for (int c = 0; c < P.n_chunks; ++c) {
                const Chunk& C = P.C[c];
                const long long lr = static_cast<long long>(t) * n_iter + last_rec[c];
                if (C.qbuf >= 0) ev[BAR_IDX_Q_EMPTY + C.qbuf].push_back({lr});
                if (C.out_dst != DST_FINAL) ev[BAR_IDX_ACT_READY + C.idx_in_layer].push_back({lr});
            }
        }
        std::vector<long long> known(64, 0);  // phases of a barrier known complete to the issuers (an earlier record waited for them)
        for (int t = 0; t < T; ++t) {
            std::vector<int> uq(2, 0);
            for (int c = 0; c < P.n_chunks; ++c) {
                const Chunk& C = P.C[c];
                for (int rr = first_rec[c]; rr <= last_rec[c]; ++rr) {
                    const long long g = static_cast<long long>(t) * n_iter + rr;
                    const uint32_t w3 = tab[4 * rr + 3];
                    auto wait = [&](int bar, uint32_t parity, long long n_req) -> bool {
                        // n_req = number of completed phases the wait is meant to see
                        if (n_req < 1 || n_req > static_cast<long long>(ev[bar].size())) { why = "wait for a phase nobody produces"; return false; }
                        if (parity != static_cast<uint32_t>((n_req - 1) & 1)) { why = "wait parity"; return false; }
                        if (known[bar] < n_req - 1) { why = "wait for phase n + 1 before phase n is known complete (stale pass)"; return false; }
                        if (ev[bar][n_req - 1].after_rec >= g) { why = "wait for an arrival that needs this record (deadlock)"; return false; }
                        if (n_req < static_cast<long long>(ev[bar].size()) && ev[bar][n_req].after_rec < g) { why = "the next phase can complete before this wait (missed phase)"; return false; }
                        known[bar] = std::max(known[bar], n_req);
                        return true;
                    };
                    const uint32_t q = w3 & 0xffu;
                    const uint32_t tpar = t & 1;
                    if (q & 3u) {
                        const int b = static_cast<int>(q & 3u) - 1;
                        if (rr != first_rec[c] || b != C.qbuf) { why = "accumulator wait on the wrong record"; return false; }
                        const long long use = static_cast<long long>(t) * 0;  // computed below from the event list
                        (void)use;
                        // this is use number u (global) of buffer b: it needs the drain of use u - 1 = event u - 1 ... i.e. u completed phases
                        long long u = 0;
                        for (int tt = 0; tt <= t; ++tt)
                            for (int cc = 0; cc < P.n_chunks; ++cc)
                                if (P.C[cc].qbuf == b && (tt < t || cc < c)) ++u;
                        const bool skip = (q & 16u) && t == 0;
                        if ((u == 0) != skip) { why = "first-use flag of an accumulator buffer"; return false; }
                        if (!skip && !wait(BAR_IDX_Q_EMPTY + b, ((q >> 2) ^ ((q >> 3) & tpar)) & 1u, u)) return false;
                    } else if (rr == first_rec[c] && C.qbuf >= 0) {
                        why = "missing accumulator wait";
                        return false;
                    }
                    for (int i = 1; i < 4; ++i) {
                        const uint32_t e = (w3 >> (8 * i)) & 0xffu;
                        if (!e) continue;
                        const int bar = static_cast<int>(e & 63u);
                        const uint32_t parity = ((e >> 6) ^ ((e >> 7) & tpar)) & 1u;
                        long long n_req;
                        if (bar == BAR_IDX_A0_READY) {
                            if (!(C.layer == 0 && C.idx_in_layer == 0 && rr == first_rec[c])) { why = "a0 wait on the wrong record"; return false; }
                            n_req = t + 1;
                        } else if (bar >= BAR_IDX_ACT_READY && bar < BAR_IDX_ACT_READY + MAX_LCHUNK) {
                            if (C.layer == 0 || C.idx_in_layer != 0) { why = "operand wait outside the first chunk of a layer"; return false; }
                            const int j = bar - BAR_IDX_ACT_READY;
                            if (j >= P.L[C.layer - 1].nchunks) { why = "operand wait for a chunk the producing layer does not have"; return false; }
                            // the k-steps of this record must not precede the awaited chunk's columns; earlier records must not reach them
                            n_req = known[bar] + 1;
                            long long cnt = 0;  // events on this barrier up to and including chunk (src_first + j) of tile t
                            for (int tt = 0; tt <= t; ++tt)
                                for (int cc = 0; cc < P.n_chunks; ++cc)
                                    if (P.C[cc].out_dst != DST_FINAL && P.C[cc].idx_in_layer == j && (tt < t || cc <= C.src_first + j)) ++cnt;
                            if (cnt != n_req) { why = "operand wait does not follow the producing chunks in order"; return false; }
                        } else {
                            why = "unknown barrier in a wait";
                            return false;
                        }
                        if (!wait(bar, parity, n_req)) return false;
                    }
                }
            }
            // every operand chunk of every layer must have been waited for by the end of the layer's first chunk: checked through
            // `known` at the end of the tile
            for (int j = 0; j < MAX_LCHUNK; ++j) {
                long long cnt = 0;
                for (int cc = 0; cc < P.n_chunks; ++cc)
                    if (P.C[cc].out_dst != DST_FINAL && P.C[cc].idx_in_layer == j) ++cnt;
                if (known[BAR_IDX_ACT_READY + j] != cnt * (t + 1)) { why = "an operand chunk is never waited for"; return false; }
            }
        }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    unsigned spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
#if VAE21_TC_WAIT_HINT
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
#else
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(static_cast<uint32_t>(VAE21_TC_WAIT_HINT))
            : "memory");
        if (ok) return;
        if ((++spins & 1023u) == 0) {  // deadlock guard: a scheduling bug must fault, not hang the GPU
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000ll) __trap();
        }
    }
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\tselp.u32 %0, 1, 0, px;\n\t}\n"
                 : "=r"(pred)
                 : "r"(0xffffffffu));
    return pred;
}
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return static_cast<uint64_t>((saddr & 0x3FFFF) >> 4) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
           (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
        "r"(ta), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
// Same with the descriptors passed as 32-bit halves (the high word -- SBO, version -- is constant).
__device__ __forceinline__ void mma_ss2(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_ts2(uint32_t d, uint32_t ta, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(d),
        "r"(ta), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
        : "memory");
}
// kind::f8f6f4 (e4m3 x e4m3, K = 32) forms of the four wrappers above / below
__device__ __forceinline__ void mma8_ss2(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %4, p;\n\t}\n" ::"r"(d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma8_ts2(uint32_t d, uint32_t ta, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(d),
        "r"(ta), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma8x2_ss2(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], da, db, %4, p;\n\t}\n" ::"r"(d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma8x2_ts2(uint32_t d, uint32_t ta, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(d),
        "r"(ta), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}


// ---- CTA-pair (cluster of 2, cta_group::2) helpers ------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(bar),
        "r"(rank)
        : "memory");
}
// (Default semantics like CUTLASS' ClusterBarrier: a cluster-scope acquire on every probe costs ~400 cycles.)
// release at cluster scope: what the arriving thread wrote / fenced before is visible to the remote waiter
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(bar),
        "r"(rank)
        : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void mma2_ss2(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma2_ts2(uint32_t d, uint32_t ta, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(d),
        "r"(ta), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
        : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void mma2_commit_both(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}\n" ::"r"(bar)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ float relu_nan(float v) {  // max that propagates NaN like np.maximum / tf.nn.relu
    float r;
    asm("max.NaN.f32 %0, %1, %2;\n" : "=f"(r) : "f"(v), "f"(0.f));
    return r;
}

// max(a, |b|, |c|) in one instruction (sm_100 three-input max; NaN operands are ignored like fmaxf)
__device__ __forceinline__ float fmax3_abs(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;\n" : "=f"(r) : "f"(a), "f"(fabsf(b)), "f"(fabsf(c)));
    return r;
}

// Split two fp32 values into packed 16-bit hi and lo words (element 0 in the low half).  The remainders v - hi are exact in fp32
// and computed as ONE packed fma (fma.rn.f32x2 on sm_100: two independent IEEE fmas, hi * (-1) + v) -- the epilogue is bound by
// its instruction count.
template <int FMT>
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const float2 v = make_float2(a, b), m1 = make_float2(-1.f, -1.f);
    if (FMT == 0) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        hi = *reinterpret_cast<uint32_t*>(&h);
        const float2 hf = make_float2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
        const float2 d = __ffma2_rn(hf, m1, v);
        __nv_bfloat162 l = __floats2bfloat162_rn(d.x, d.y);
        lo = *reinterpret_cast<uint32_t*>(&l);
    } else {
        __half2 h = __floats2half2_rn(a, b);
        hi = *reinterpret_cast<uint32_t*>(&h);
        const float2 d = __ffma2_rn(__half22float2(h), m1, v);
        __half2 l = __floats2half2_rn(d.x, d.y);
        lo = *reinterpret_cast<uint32_t*>(&l);
    }
}

// Four fp32 values -> the operand words of one k-quad.  FMT 0/1: two hi words + two lo words (16-bit pairs).
// FMT 2: two fp16 hi words, ONE word of four e4m3(v) and ONE word of four e4m3((v - hi) * 2^11).
template <int FMT>
__device__ __forceinline__ void split4_f8(float v0, float v1, float v2, float v3, uint32_t& h01, uint32_t& h23, uint32_t& b_hi8,
                                          uint32_t& b_lo8) {
    const __half2 ha = __floats2half2_rn(v0, v1), hb = __floats2half2_rn(v2, v3);
    h01 = *reinterpret_cast<const uint32_t*>(&ha);
    h23 = *reinterpret_cast<const uint32_t*>(&hb);
    const uint32_t p0 = __nv_cvt_float2_to_fp8x2(make_float2(v0, v1), __NV_SATFINITE, __NV_E4M3);
    const uint32_t p1 = __nv_cvt_float2_to_fp8x2(make_float2(v2, v3), __NV_SATFINITE, __NV_E4M3);
    b_hi8 = p0 | (p1 << 16);
    // (v - hi) * 2^11: both steps exact in fp32, two packed instructions per pair
    const float2 m1 = make_float2(-1.f, -1.f), sc = make_float2(A_LO_SCALE, A_LO_SCALE);
    const float2 da = __fmul2_rn(__ffma2_rn(__half22float2(ha), m1, make_float2(v0, v1)), sc);
    const float2 db = __fmul2_rn(__ffma2_rn(__half22float2(hb), m1, make_float2(v2, v3)), sc);
    const uint32_t q0 = __nv_cvt_float2_to_fp8x2(da, __NV_SATFINITE, __NV_E4M3);
    const uint32_t q1 = __nv_cvt_float2_to_fp8x2(db, __NV_SATFINITE, __NV_E4M3);
    b_lo8 = q0 | (q1 << 16);
}
// 16 consecutive features of one row -> the 16 operand words of a k-step: w[0..7] = 16-bit hi pairs (k-groups 0 and 1),
// w[8..15] = second tile (FMT 0/1: lo pairs; FMT 2: w[8..11] e4m3(v) for the 16 k, w[12..15] e4m3 of the scaled remainders).
template <int FMT>
__device__ __forceinline__ void split16(const float (&v)[16], uint32_t (&w)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (FMT == 2) {
            split4_f8<FMT>(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3], w[2 * q], w[2 * q + 1], w[8 + q], w[12 + q]);
        } else {
            split2<FMT>(v[4 * q], v[4 * q + 1], w[2 * q], w[8 + 2 * q]);
            split2<FMT>(v[4 * q + 2], v[4 * q + 3], w[2 * q + 1], w[8 + 2 * q + 1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------
// One k-step of one accumulator chunk: the kind::f16 hi x hi product plus the correction MMA(s) of the operand format.
//   TS (A operand in TMEM):  a = TMEM column of the 16-bit hi pairs, a2 = a + 8 (second operand part)
//   SS (A operand in smem):  a = low descriptor word of the hi tile, a2 = of the second tile
//   b / b2: low descriptor words of the weight hi tile / second tile; `hi` = constant high descriptor word
template <int FMT, bool PAIR>
__device__ __forceinline__ void kstep_mma(bool ts, uint32_t d, uint32_t a, uint32_t a2, uint32_t b, uint32_t b2, uint32_t hi, uint32_t idesc,
                                          uint32_t acc0) {
    if (ts) {
        if (FMT == 2) {
            if (PAIR) { mma2_ts2(d, a, b, hi, idesc, acc0); mma8x2_ts2(d, a2, b2, hi, idesc, 1u); }
            else { mma_ts2(d, a, b, hi, idesc, acc0); mma8_ts2(d, a2, b2, hi, idesc, 1u); }
        } else if (PAIR) {
            mma2_ts2(d, a, b, hi, idesc, acc0); mma2_ts2(d, a, b2, hi, idesc, 1u); mma2_ts2(d, a2, b, hi, idesc, 1u);
        } else {
            mma_ts2(d, a, b, hi, idesc, acc0); mma_ts2(d, a, b2, hi, idesc, 1u); mma_ts2(d, a2, b, hi, idesc, 1u);
        }
    } else {
        if (FMT == 2) {
            if (PAIR) { mma2_ss2(d, a, b, hi, idesc, acc0); mma8x2_ss2(d, a2, b2, hi, idesc, 1u); }
            else { mma_ss2(d, a, b, hi, idesc, acc0); mma8_ss2(d, a2, b2, hi, idesc, 1u); }
        } else if (PAIR) {
            mma2_ss2(d, a, b, hi, idesc, acc0); mma2_ss2(d, a, b2, hi, idesc, 1u); mma2_ss2(d, a2, b, hi, idesc, 1u);
        } else {
            mma_ss2(d, a, b, hi, idesc, acc0); mma_ss2(d, a, b2, hi, idesc, 1u); mma_ss2(d, a2, b, hi, idesc, 1u);
        }
    }
}

// The parameter transform of one row (preprocess.py:74-78, :105-108): floor substitution and log10 in fp64 like the reference and the
// FP32 path, then ONE fp64 fma for the affine map x = (t - pmin) * (2 / prange) - 1 (the reference divides: the two differ by at most
// one fp64 rounding, far below the fp32 cast that follows and irrelevant at the tensor-core tolerance).  Out of line: it contains
// the double-precision log10 and runs in the two prologue warps only.
__device__ __noinline__ void prologue_row(const LaunchArgs& a, const NormConsts& nc, long long grow, int K0, float* x /*[16]*/) {
#pragma unroll 1
    for (int j = 0; j < 16; ++j) x[j] = 0.f;
    if (grow >= a.n) return;
    if (a.in_mode == IN_GRID) {
        grid_point(a, static_cast<unsigned long long>(a.row_base + grow), K0, x);
        return;
    }
    if (a.in_mode == IN_NORMALISED_F32) {
#pragma unroll 1
        for (int j = 0; j < K0; ++j) x[j] = reinterpret_cast<const float*>(a.in)[grow * K0 + j];
        return;
    }
    const bool f32_in = (a.in_mode == IN_PARAMS_F32);
    // all raw parameters of the row first: independent loads, ONE memory latency per row instead of one per parameter
    double pv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (j < K0)
            pv[j] = f32_in ? static_cast<double>(reinterpret_cast<const float*>(a.in)[grow * K0 + j]) : reinterpret_cast<const double*>(a.in)[grow * K0 + j];
    }
#pragma unroll 1
    for (int j = 0; j < K0; ++j) {
        double p = pv[j];
        if (j == nc.floor_col && p == 0.0) p = f32_in ? static_cast<double>(static_cast<float>(nc.floor_val)) : nc.floor_val;
        double t = p;
        if (nc.log_mask[j]) {
            t = log10(p);
            if (f32_in) t = static_cast<double>(static_cast<float>(t));  // numpy takes the log of a float32 array in float32
        }
        x[j] = static_cast<float>(fma(t - nc.pmin[j], nc.pscale[j], -1.0));
    }
}

// The same for a compile-time parameter count (the emulators have 7): everything unrolled, so the normalisation constants are
// immediate constant-bank operands and the row's values stay in registers -- the generic version above indexes the constant bank and
// a local array dynamically, which made one 32-row round cost ~14k cycles (round-2 measurement) and the lone prologue warp the
// kernel's bottleneck.  The double-precision log10 stays one out-of-line copy.
__device__ __noinline__ double log10_f64(double p) { return log10(p); }

template <int K0>
__device__ __forceinline__ void prologue_row_fixed(const LaunchArgs& a, const NormConsts& nc, long long grow, float (&x)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = 0.f;
    if (grow >= a.n) return;
    if (a.in_mode == IN_NORMALISED_F32) {
#pragma unroll
        for (int j = 0; j < K0; ++j) x[j] = reinterpret_cast<const float*>(a.in)[grow * K0 + j];
        return;
    }
    const bool f32_in = (a.in_mode == IN_PARAMS_F32);
    double pv[K0];
    if (f32_in) {
#pragma unroll
        for (int j = 0; j < K0; ++j) pv[j] = static_cast<double>(reinterpret_cast<const float*>(a.in)[grow * K0 + j]);
    } else {
#pragma unroll
        for (int j = 0; j < K0; ++j) pv[j] = reinterpret_cast<const double*>(a.in)[grow * K0 + j];
    }
#pragma unroll
    for (int j = 0; j < K0; ++j) {
        double p = pv[j];
        if (j == nc.floor_col && p == 0.0) p = f32_in ? static_cast<double>(static_cast<float>(nc.floor_val)) : nc.floor_val;
        double t = p;
        if (nc.log_mask[j]) {
            t = log10_f64(p);
            if (f32_in) t = static_cast<double>(static_cast<float>(t));  // numpy takes the log of a float32 array in float32
        }
        x[j] = static_cast<float>(fma(t - nc.pmin[j], nc.pscale[j], -1.0));
    }
}

// Tensor maps of the output for the OM_ROWS store path (encoded per launch on the host, see make_store_maps): the (n, n_out) float32
// output viewed as [n / 4 super-rows][4 n_out words] -- a row pitch of 4 n_out bytes x 4 is a multiple of 16 bytes for ANY n_out,
// which a tensor map requires and the 1804-byte rows of the 451-bin spectra are not -- one map per distinct box width.
struct alignas(64) StoreMaps {
    CUtensorMap m[4];
    long long rows4;  // 4 * (n / 4): rows covered by complete super-rows (0: tensor stores off for this launch)
};

// Output modes as a template parameter (each instantiation carries only its own final-layer code):
enum : int { OM_ROWS = 0 /* OUT_PREDICT / OUT_NORMALISED: spectra written */, OM_CHI2 = 1, OM_ERROR = 2 };

// hidden activations beyond this magnitude leave the range of the 8-bit / 16-bit operand parts (counted, see LaunchArgs::sat)
template <int FMT>
__device__ __forceinline__ constexpr float sat_limit() { return FMT == 2 ? 480.f : 65504.f; }

// CG = 1: one CTA per 128-row tile (cta_group::1).  CG = 2: a cluster of two CTAs works on a 256-row
// super-tile with cta_group::2 MMAs (M = 256): rank 0 issues every MMA for both CTAs, each CTA streams
// HALF of every weight tile into its own shared memory and runs its own epilogue on its own 128 rows.
template <int FMT, int CG, int OM>  // FMT 0: bf16 split (3 MMAs per k-step), 1: fp16 split (3), 2: fp16 + e4m3 corrections (2)
__global__ void __launch_bounds__(NTHREADS, 1)
vae21_tc_kernel(const __grid_constant__ Plan P, const __grid_constant__ NormConsts nc, const __grid_constant__ LaunchArgs a,
                const uint8_t* __restrict__ wimg, const float* __restrict__ bias_g, const __grid_constant__ StoreMaps smaps) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 127u) & ~127u;  // shared-window address of the carve-up
    uint8_t* sm = smem_raw + (base - raw);

    // `warp` is the ROLE index (0 producer, 1 MMA, 2-3 prologue, 4.. epilogue); the epilogue role 4 + e runs on hardware
    // warp e (VAE21_TC_CTRL_LAST, measured neutral), so its TMEM sub-partition (hardware warp id mod 4) is still `warp & 3`.
    // The hardware warp index goes through a shuffle so that the compiler can PROVE it warp-uniform: the role branches below are then
    // uniform branches, and the loop counters, chunk records and MMA descriptors of the control warps live in uniform registers
    // instead of being moved there (R2UR) before every tcgen05 instruction.
    const int tid = threadIdx.x, lane = tid & 31;
    const int hw_warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
#if VAE21_TC_CTRL_LAST
    const int warp = hw_warp < NEPI / 32 ? hw_warp + 4 : hw_warp - NEPI / 32;
#else
    const int warp = hw_warp;
#endif
    constexpr bool PAIR = (CG == 2);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = (rank == 0);
    // work units: 128-row tiles (CG 1) or 256-row super-tiles (CG 2); this CTA's tile of unit u is CG*u + rank
    const long long nunits = (a.n + CG * MT - 1) / (CG * MT);
    const long long unit0 = blockIdx.x / CG, ustep = gridDim.x / CG;

    // barrier addresses
    const uint32_t bar0 = base + P.off_bar;
    auto bar_ring_full = [&](int s) { return bar0 + 8u * s; };
    auto bar_ring_empty = [&](int s) { return bar0 + 8u * (MAX_SLOTS + s); };
    auto bar_chunk_full = [&](int i) { return bar0 + 8u * (2 * MAX_SLOTS + i); };
    auto bar_q_empty = [&](int b) { return bar0 + 8u * (BAR_IDX_Q_EMPTY + b); };
    auto bar_act_ready = [&](int j) { return bar0 + 8u * (BAR_IDX_ACT_READY + j); };  // j < MAX_LCHUNK
    const uint32_t bar_a0_ready = bar0 + 8u * BAR_IDX_A0_READY;
    const uint32_t bar_a0_free = bar0 + 8u * BAR_IDX_A0_FREE;  // layer-0 MMAs of the tile have read a0
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + P.off_bar + 8 * (2 * MAX_SLOTS + NFULL + 4 + MAX_LCHUNK));
    float* s_chi = reinterpret_cast<float*>(sm + P.off_bar + BAR_BYTES) + 32;  // [2][EPS-1][128] chi^2 partials of the other column shares
    float* s_amp = s_chi + 2 * (EPS - 1) * 128;                                 // [2][EPS-1][128] |truth| maxima (OM_ERROR)

    float* s_bias = reinterpret_cast<float*>(sm + P.off_bias);
    float* s_s0 = reinterpret_cast<float*>(sm + P.off_s0);
    float* s_obs = reinterpret_cast<float*>(sm + P.off_obs);
    float* s_isig = reinterpret_cast<float*>(sm + P.off_isig);

    // ---- one-time setup -------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < (PAIR ? P.nslots2 : P.nslots); ++s) {
            // pair leader: its own producer (arrive + tx bytes) and the peer's forwarded "my half landed"
            mbar_init(bar_ring_full(s), (PAIR && leader) ? 2 : 1);
            mbar_init(bar_ring_empty(s), 1);
        }
        // epilogue -> MMA hand-offs: ONE arrival per epilogue warp (elected lane after __syncwarp).  In a pair the follower's epilogue
        // warps arrive DIRECTLY on the leader's barriers (remote arrive), so those count both CTAs.
        const uint32_t both = (PAIR && leader) ? 2u : 1u;
        for (int i = 0; i < NFULL; ++i) mbar_init(bar_chunk_full(i), 2);  // the commits behind the last two ring slots of a chunk
        for (int b = 0; b < 2; ++b) mbar_init(bar_q_empty(b), (NEPI / 32) * both);
        for (int j = 0; j < MAX_LCHUNK; ++j) mbar_init(bar_act_ready(j), (NEPI / 32) * both);
        mbar_init(bar_a0_ready, both);  // one arrival from the prologue warp (of each CTA)
        mbar_init(bar_a0_free, static_cast<uint32_t>(P.l0_iters[PAIR ? 1 : 0]));  // one commit behind every ring slot of layer 0
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    {
        const Layer& LL = P.L[P.n_layers - 1];
        const int nop = LL.Npad;
        for (int i = tid; i < P.bias_smem; i += NTHREADS) s_bias[i] = bias_g[i];
        {  // this kernel variant's issue table (raw words behind the biases)
            const uint32_t* src = reinterpret_cast<const uint32_t*>(bias_g) + P.iter_off[PAIR ? 1 : 0];
            uint32_t* dst = reinterpret_cast<uint32_t*>(sm + P.off_iter);
            for (int i = tid; i < 4 * P.n_iter[PAIR ? 1 : 0]; i += NTHREADS) dst[i] = src[i];
        }
        for (int n = tid; n < nop; n += NTHREADS) {
            const float b = bias_g[LL.bias_off + n];
            float s0 = b, ob = 0.f, is = 0.f;
            if (n < P.n_out) {
                if (a.out_mode != OUT_NORMALISED) s0 = fmaf(b, nc.sd, a.mu[n]);
                if (OM == OM_CHI2) {
                    ob = a.obs[n];
                    is = a.isig[n];
                }
                if (OM == OM_ERROR) is = a.isig[n];  // 1 inside the frequency band, 0 outside
            }
            s_s0[n] = s0;
            s_obs[n] = ob;
            s_isig[n] = is;
        }
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // both CTAs: barriers initialised and TMEM allocated before any cross-CTA traffic
    tc_fence_after();
    const uint32_t tm = *tmem_slot;
    const int n_chunks = P.n_chunks, nslots = PAIR ? P.nslots2 : P.nslots;
    const uint32_t ring0 = base + (PAIR ? P.off_ring2 : P.off_ring), slot_bytes = static_cast<uint32_t>(PAIR ? P.slot_bytes2 : P.slot_bytes);

    if (warp == 0) {
        // ===================== producer: stream the weight image through the ring ============
        if (lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            const uint8_t* img = PAIR ? wimg + P.w_bytes + static_cast<size_t>(rank) * (P.w_bytes / 2) : wimg;
#pragma unroll 1
            for (long long unit = unit0; unit < nunits; unit += ustep) {
#pragma unroll 1
                for (int c = 0; c < n_chunks; ++c) {
                    const Chunk& C = P.C[c];
                    const uint32_t bytes = static_cast<uint32_t>(C.ncols) * (PAIR ? 32u : 64u);  // one k-step of this CTA's rows
                    const uint8_t* src = img + (PAIR ? C.w_off / 2 : C.w_off);
                    const int nst = C.nstages, kps = PAIR ? C.kps2 : 1;
#pragma unroll 1
                    for (int s = 0; s < nst; s += kps) {
                        const uint32_t sbytes = bytes * static_cast<uint32_t>(min(kps, nst - s));  // k-steps in this slot
                        mbar_wait(bar_ring_empty(slot), phase ^ 1u);
                        mbar_expect_tx(bar_ring_full(slot), sbytes);
                        asm volatile(
                            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                                ring0 + slot * slot_bytes),
                            "l"(src), "r"(sbytes), "r"(bar_ring_full(slot))
                            : "memory");
                        src += sbytes;
                        if (++slot == nslots) {
                            slot = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && PAIR && !leader) {
        // ===================== follower of a pair: relay "my half of this ring slot has landed" to the issuers =========
        // (a bulk copy cannot signal a barrier in another CTA; the leader's MMAs read both CTAs' halves)
        int slot = 0;
        uint32_t rphase = 0;
        const int n_iter = P.n_iter[1];
#pragma unroll 1
        for (long long unit = unit0; unit < nunits; unit += ustep) {
#pragma unroll 1
            for (int it = 0; it < n_iter; ++it) {
                mbar_wait(bar_ring_full(slot), rphase);
                if (lane == 0) mbar_arrive_remote(bar_ring_full(slot), 0);
                __syncwarp();
                if (++slot == nslots) {
                    slot = 0;
                    rphase ^= 1u;
                }
            }
        }
    } else if ((warp == 1 || warp == WARP_ISSUER_B) && (!PAIR || leader)) {
        // ===================== MMA issuers ===================================================
        // The issuers interpret the tile's issue table (build_iters): one record per ring slot.  With two issuers (the default), warp
        // `me` takes the records with global index g = me (mod 2) and a token -- two named barriers used as a baton, with
        // tcgen05.fence on both sides -- orders the MMAs of record g behind those of record g - 1: every tcgen05.mma of the CTA
        // (pair) is issued in table order, so the accumulation order, hence every output bit, is fixed, while each warp's barrier
        // waits, record decoding and descriptor set-up overlap the other warp's issue.  The whole warp runs the (warp-uniform) loop;
        // only the tcgen05 instructions are predicated on one elected lane.
        const int me = (warp == 1) ? 0 : 1;
        const int nis = P.issuers;
        const int n_iter = P.n_iter[PAIR ? 1 : 0];
        if (me < nis && unit0 < nunits) {
            const uint4* tab = reinterpret_cast<const uint4*>(sm + P.off_iter);
            const uint32_t fmtbits = (FMT == 0) ? 1u : 0u;
            const uint32_t idesc_base = (1u << 4) | (fmtbits << 7) | (fmtbits << 10) | ((static_cast<uint32_t>(CG * 128) >> 4) << 24);
            // descriptor high words are constant: SBO = 128 B, version 1; LBO goes into the low word
            const uint32_t desc_hi = (128u >> 4) | (1u << 14);
            const uint32_t a_base32 = ((base & 0x3FFFFu) >> 4) | ((static_cast<uint32_t>(A_KG_BYTES) >> 4) << 16);
            const uint32_t ring16 = (ring0 & 0x3FFFFu) >> 4, slot16 = slot_bytes >> 4;
#if VAE21_TC_TIMING
            long long tm_total = clock64(), tm_flag = 0, tm_ring = 0, tm_token = 0, tm_issue = 0;
#define TSTART const long long t_s_ = clock64();
#define TADD(var) var += clock64() - t_s_;
#else
#define TSTART
#define TADD(var)
#endif
            long long unit = unit0;
            const uint32_t last_owner = static_cast<uint32_t>((((nunits - unit0 + ustep - 1) / ustep) * n_iter - 1) & 1);
            uint32_t tpar = 0;      // parity of the CTA's tile count
            uint32_t first_tile = 1;
            uint32_t seq_base = 0;  // chunks of earlier tiles (index of the chunk_full ring)
            int it = me, slot = me;  // n_iter >= 2, nslots >= 2
            uint32_t rphase = 0;
            bool started = false;   // this warp has passed at least one token wait / the other has issued before
            uint4 rec = tab[it];
#pragma unroll 1
            while (true) {
                // next record (prefetched: the load overlaps this iteration's waits)
                int it_n = it + nis;
                const bool roll = it_n >= n_iter;
                if (roll) it_n -= n_iter;
                const uint4 nxt = tab[it_n];

                const uint32_t w2 = rec.z, w3 = rec.w;
                const uint32_t ncols = w2 & 0x1ffu;
                const int nk = static_cast<int>((w2 >> 9) & 7u);
                const uint32_t full = bar_ring_full(slot);
                {
                    TSTART
                    if (!mbar_try(full, rphase)) mbar_wait(full, rphase);
                    TADD(tm_ring)
#if VAE21_TC_TIMING
                    if (lane == 0 && it < 256) atomicAdd(&g_tc_rec[1][it], static_cast<unsigned long long>(clock64() - t_s_));
#endif
                }
                if (nis == 2 && (started || me == 1)) {  // the baton: the other warp has issued the previous record
                    TSTART
                    if (me == 0) asm volatile("bar.sync 3, 64;\n" ::: "memory");
                    else asm volatile("bar.sync 2, 64;\n" ::: "memory");
                    TADD(tm_token)
                }
                // Accumulator buffer drained / operands converted (at most a few records per chunk).  These waits come AFTER the baton:
                // mbarrier waits are by phase parity, and a wait for phase n + 1 only means something once phase n is known to be over
                // -- which the owner of an earlier record has waited for before issuing it.
                if (w3) {
                    TSTART
                    const uint32_t q = w3 & 0xffu;
                    if ((q & 3u) && !((q & 16u) && first_tile))
                        mbar_wait(bar0 + 8u * (BAR_IDX_Q_EMPTY + (q & 3u) - 1u), ((q >> 2) ^ ((q >> 3) & tpar)) & 1u);
#pragma unroll
                    for (int i = 1; i < 4; ++i) {
                        const uint32_t e = (w3 >> (8 * i)) & 0xffu;
                        if (e) mbar_wait(bar0 + 8u * (e & 63u), ((e >> 6) ^ ((e >> 7) & tpar)) & 1u);
                    }
                    TADD(tm_flag)
#if VAE21_TC_TIMING
                    if (lane == 0 && it < 256) atomicAdd(&g_tc_rec[0][it], static_cast<unsigned long long>(clock64() - t_s_));
#endif
                }
#if VAE21_TC_TIMING
                if (lane == 0 && it < 256) atomicAdd(&g_tc_rec[2][it], 1ull);
#endif
                started = true;
                tc_fence_after();
                {
                    TSTART
                    if (elect_one()) {
                        const uint32_t idesc = idesc_base | rec.y;
                        const uint32_t d = tm + (rec.x >> 16);
                        const bool ts = (w2 & IT_TS) != 0;
                        const uint32_t b_kg16 = ncols / CG;                     // bytes between B k-groups (this CTA's rows) / 16
                        const uint32_t bj = ring16 + slot * slot16 + (b_kg16 << 16);  // low descriptor word of the slot's first hi tile
                        const uint32_t b_lo16 = 2u * b_kg16, kstep16 = 4u * b_kg16;    // hi tile -> second tile, k-step -> k-step (16 B units)
                        const uint32_t acur = ts ? tm + (rec.x & 0xffffu) : a_base32 + (rec.x & 0xffffu);
                        const uint32_t astep = ts ? 16u : static_cast<uint32_t>(KSTEP_BYTES >> 4);
                        const uint32_t a2off = ts ? 8u : ((2u * A_KG_BYTES) >> 4);
                        const uint32_t acc0 = (w2 & IT_FIRST) ? 0u : 1u;
                        // up to 4 k-steps per slot, fully unrolled
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (j < nk && !(DBG & 1))
                                kstep_mma<FMT, PAIR>(ts, d, acur + j * astep, acur + j * astep + a2off, bj + j * kstep16, bj + j * kstep16 + b_lo16,
                                                     desc_hi, idesc, j > 0 ? 1u : acc0);
                        }
                        // the slot is freed -- in both CTAs of a pair -- by a commit behind the MMAs of its last k-step
                        if (PAIR) mma2_commit_both(full + 8u * MAX_SLOTS);
                        else mma_commit(full + 8u * MAX_SLOTS);
                        const uint32_t ncommit = (w2 >> 14) & 3u;
                        if (ncommit) {
                            const uint32_t cf = bar_chunk_full((seq_base + (w2 >> 20)) & (NFULL - 1));
                            if (PAIR) mma2_commit_both(cf);
                            else mma_commit(cf);
                            if (ncommit == 2) {
                                if (PAIR) mma2_commit_both(cf);
                                else mma_commit(cf);
                            }
                        }
                        if (w2 & IT_A0FREE) {  // the prologue warp may overwrite a0 once the layer-0 MMAs are done
                            if (PAIR) mma2_commit_both(bar_a0_free);
                            else mma_commit(bar_a0_free);
                        }
                    }
                    __syncwarp();
                    TADD(tm_issue)
                }
                if (nis == 2) {  // pass the baton
                    tc_fence_before();
                    if (me == 0) asm volatile("bar.arrive 2, 64;\n" ::: "memory");
                    else asm volatile("bar.arrive 3, 64;\n" ::: "memory");
                }
                // advance to this warp's next record
                slot += nis;
                if (slot >= nslots) {
                    slot -= nslots;
                    rphase ^= 1u;
                }
                it = it_n;
                rec = nxt;
                if (roll) {
                    unit += ustep;
                    if (unit >= nunits) break;
                    tpar ^= 1u;
                    first_tile = 0;
                    seq_base += static_cast<uint32_t>(n_chunks);
                }
            }
            // the warp that did not issue the CTA's last record takes the last baton, so that no barrier is left half-arrived
            if (nis == 2 && static_cast<uint32_t>(me) != last_owner) {
                if (me == 0) asm volatile("bar.sync 3, 64;\n" ::: "memory");
                else asm volatile("bar.sync 2, 64;\n" ::: "memory");
            }
#if VAE21_TC_TIMING
            if (lane == 0 && blockIdx.x < 160) {
                long long* o = g_tc_timing[blockIdx.x] + 8 * me;
                o[0] = clock64() - tm_total;
                o[1] = tm_flag; o[2] = tm_ring; o[3] = tm_token; o[4] = tm_issue;
            }
#endif
        }
    } else if (warp == 2) {
        // ===================== prologue warp: layer-0 operand of every tile ====================
        // fused parameter transform (preprocess.py:74-78, :105-108) -> a0 (k padded to 16), four rounds of 32 rows.  With at most 8
        // input parameters (the emulators have 7) the operand words of the NEXT tile are computed first and held in registers (8
        // words per row: only the first k-group of a0 is ever non-zero), and only then does the warp wait for the current tile's
        // layer-0 MMAs to release a0 -- parameter loads and fp64 logarithms are off every critical path.  Wider inputs: after the wait.
        const int K0 = P.K0;
        uint8_t* const a0 = sm + P.off_a0;
        for (int i = lane; i < KSTEP_BYTES / 16; i += 32) reinterpret_cast<uint4*>(a0)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        uint32_t nfree = 0;
#pragma unroll 1
        for (long long unit = unit0; unit < nunits; unit += ustep, ++nfree) {
            const long long tile = CG * unit + rank;
#if VAE21_TC_TIMING
            const long long tp0 = clock64();
#endif
            if (K0 <= 8) {
                uint32_t w8[MT / 32][8];
#pragma unroll
                for (int rr = 0; rr < MT / 32; ++rr) {
                    float x[16];
                    if (K0 == 7 && a.in_mode != IN_GRID) prologue_row_fixed<7>(a, nc, tile * MT + rr * 32 + lane, x);
                    else prologue_row(a, nc, tile * MT + rr * 32 + lane, K0, x);
                    uint32_t w[16];
                    split16<FMT>(x, w);
                    w8[rr][0] = w[0]; w8[rr][1] = w[1]; w8[rr][2] = w[2]; w8[rr][3] = w[3];
                    if (FMT == 2) { w8[rr][4] = w[8]; w8[rr][5] = w[9]; w8[rr][6] = w[12]; w8[rr][7] = w[13]; }
                    else { w8[rr][4] = w[8]; w8[rr][5] = w[9]; w8[rr][6] = w[10]; w8[rr][7] = w[11]; }
                }
#if VAE21_TC_TIMING
                const long long tp1 = clock64();
#endif
                if (nfree > 0) mbar_wait(bar_a0_free, (nfree - 1u) & 1u);
#if VAE21_TC_TIMING
                if (lane == 0 && blockIdx.x < 160) {
                    g_tc_timing[blockIdx.x][5] += tp1 - tp0;        // operand computation
                    g_tc_timing[blockIdx.x][6] += clock64() - tp1;  // wait for a0 to be released
                }
#endif
#pragma unroll
                for (int rr = 0; rr < MT / 32; ++rr) {
                    const int row = rr * 32 + lane;
                    *reinterpret_cast<uint4*>(a0 + row * 16) = make_uint4(w8[rr][0], w8[rr][1], w8[rr][2], w8[rr][3]);
                    if (FMT == 2) {  // second tile: k-group 0 = e4m3(x) of the 16 k, k-group 1 = e4m3 of the scaled remainders
                        *reinterpret_cast<uint2*>(a0 + 2 * A_KG_BYTES + row * 16) = make_uint2(w8[rr][4], w8[rr][5]);
                        *reinterpret_cast<uint2*>(a0 + 3 * A_KG_BYTES + row * 16) = make_uint2(w8[rr][6], w8[rr][7]);
                    } else {
                        *reinterpret_cast<uint4*>(a0 + 2 * A_KG_BYTES + row * 16) = make_uint4(w8[rr][4], w8[rr][5], w8[rr][6], w8[rr][7]);
                    }
                }
            } else {
                if (nfree > 0) mbar_wait(bar_a0_free, (nfree - 1u) & 1u);
#pragma unroll 1
                for (int rr = 0; rr < MT / 32; ++rr) {
                    const int row = rr * 32 + lane;
                    float x[16];
                    prologue_row(a, nc, tile * MT + row, K0, x);
                    uint32_t w[16];
                    split16<FMT>(x, w);
                    *reinterpret_cast<uint4*>(a0 + row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(a0 + A_KG_BYTES + row * 16) = make_uint4(w[4], w[5], w[6], w[7]);
                    *reinterpret_cast<uint4*>(a0 + 2 * A_KG_BYTES + row * 16) = make_uint4(w[8], w[9], w[10], w[11]);
                    *reinterpret_cast<uint4*>(a0 + 3 * A_KG_BYTES + row * 16) = make_uint4(w[12], w[13], w[14], w[15]);
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (PAIR && !leader) mbar_arrive_remote(bar_a0_ready, 0);
                else mbar_arrive(bar_a0_ready);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps ================================================
        // thread = tile row = TMEM lane; the EPS warps that share a TMEM sub-partition split the 16-column groups of every
        // accumulator chunk into EPS contiguous ranges (warp share h takes groups [ng h / EPS, ng (h + 1) / EPS)).
        const int ew = warp - 4;
        const int half = ew >> 2;                 // which share of the 16-column groups
        const int sub = warp & 3;                 // TMEM sub-partition this warp may access
        const int row = sub * 32 + lane;          // tile row == TMEM lane
        const uint32_t tlane = static_cast<uint32_t>(sub * 32) << 16;
        const int NO = P.n_out;
        uint32_t seq = 0;
        uint32_t tcount = 0;
        uint32_t fin_cnt = 0;  // final-layer chunks staged so far (staging buffer parity)
        float vmax = 0.f;  // largest hidden activation this thread converted (operand-range check, FMT 1 / 2)
        // OM_ROWS output path.  The 1804-byte rows of the (n, 451) output rule out a row-pitched tensor map (pitch must be a multiple
        // of 16 bytes) and make plain stores expensive (round-2 ablation: the transposes + 64-byte-segment STGs of round 1 cost
        // 0.31 ms of a 1.73 ms launch, one 1-D bulk store per thread and chunk 0.45 ms).  But FOUR rows are 7216 bytes: the output is
        // a [n / 4][4 n_out] tensor whose "super-row" r4 holds rows 4 r4 .. 4 r4 + 3, and the chunk's columns of all rows with the
        // same residue q = row mod 4 form a box -- if the box starts at a 16-byte boundary, i.e. if rows of residue q start `shift`
        // = (-q n_out) mod 4 columns into the chunk.  So every thread stages the columns [shift, shift + W) of its row's chunk piece
        // in the dense [8 super-rows][W] tile of its (sub-partition, residue) -- W = 12 (mod 16) makes the 32 rows of a warp hit 32
        // different banks --, writes the <= 3 + 4 columns around it with scalar stores straight from registers, and after one
        // named barrier of the sub-partition's four warps FOUR lanes issue the four box stores (cp.async.bulk.tensor.2d): 16 store
        // instructions per chunk and CTA instead of 2048.  Two staging buffers alternate between final chunks; the issuing lanes wait
        // for the boxes of earlier chunks to have been read before they enter the next barrier, which frees the older buffer for
        // everybody.  Rows beyond the last complete super-row, launches with an unaligned output and chunks without a box are
        // written with plain stores.
        const int q_res = row & 3;                                    // row residue (tile bases are multiples of 4)
        const int shift = (4 - ((q_res * NO) & 3)) & 3;
        const bool tma_launch = (OM == OM_ROWS) && smaps.rows4 > 0;
        uint8_t* const stage_base = sm + P.off_stage + (sub * 4 + q_res) * P.stage_sq;
        const bool issuer = (half == 0 && lane < 4);                  // lane q of the first column share issues the box of residue q
                                                                      // (one lane of each of the four shares: measured 2 % slower)

        // hand-off to the MMA issuers, which live in the leader CTA of a pair
        auto signal_mma = [&](uint32_t bar) {
            if (PAIR && !leader) mbar_arrive_remote(bar, 0);
            else mbar_arrive(bar);
        };

#pragma unroll 1
        for (long long unit = unit0; unit < nunits; unit += ustep, ++tcount) {
            const long long tile = CG * unit + rank;
            const long long grow = tile * MT + row;
            float chi = 0.f, amp = 0.f;
#pragma unroll 1
            for (int c = 0; c < n_chunks; ++c) {
                // The chunk record is read BEFORE the wait: the (dynamically indexed) constant loads then overlap the wait.
                const Chunk& C = P.C[c];
                const int c_bias_n0 = C.bias_n0, c_ncols = C.ncols, c_dcol = C.dcol, c_n0 = C.n0, c_qbuf = C.qbuf, c_idx = C.idx_in_layer;
                const int out_dst = C.out_dst;
                const bool do_relu = C.relu != 0;
                const float inv_s8 = (FMT == 2) ? C.inv_s8 : 1.f;
                mbar_wait(bar_chunk_full(seq & (NFULL - 1)), (seq / NFULL) & 1u);
                ++seq;
                tc_fence_after();
                if (OM == OM_ROWS && c == 0 && tcount > 0) {
                    // the output staging buffers alias the activation buffer: before this tile's first activations are written, the box
                    // stores of the previous tile must have read them
                    if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
                    asm volatile("bar.sync 1, %0;\n" ::"n"(NEPI) : "memory");
                }
                const float* bl = s_bias + c_bias_n0;
                const int ng = c_ncols / 16;
                const int g_lo = (ng * half) / EPS, g_hi = (ng * (half + 1)) / EPS;
                const uint32_t tbase = tm + tlane + static_cast<uint32_t>(c_dcol);
                // this chunk's tensor-store box (final layer, OM_ROWS): width st_w, this thread's row of the staging tile
                const int st_w = (OM == OM_ROWS && out_dst == DST_FINAL && tma_launch) ? C.st_w : 0;
                const bool staged = st_w > 0 && grow < smaps.rows4;  // else: plain stores
                float* sstage = nullptr;  // word `col` of the chunk goes to sstage[col] for shift <= col < shift + st_w
                if (OM == OM_ROWS && out_dst == DST_FINAL && tma_launch && P.stage_bufs == 1) {
                    // a single staging buffer: the previous chunk's boxes must have been read before anybody overwrites it
                    if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
                    asm volatile("bar.sync %0, 128;\n" ::"r"(4 + sub) : "memory");
                }
                if (OM == OM_ROWS && out_dst == DST_FINAL) {
                    sstage = reinterpret_cast<float*>(stage_base + ((P.stage_bufs == 2) ? (fin_cnt & 1u) * (16 * P.stage_sq) : 0)) + (lane >> 2) * st_w - shift;
                    ++fin_cnt;
                }
                // group body (instantiated twice: the accumulator reads are software-pipelined over two register sets, so the load of
                // this warp's next group is in flight while the current one is converted)
                auto process = [&](uint32_t (&r)[16], int g) {
                    float v[16];
                    if (out_dst != DST_FINAL) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 b4 = *reinterpret_cast<const float4*>(bl + 16 * g + 4 * q);
                            if (FMT == 2) {  // accumulators carry the weight scale S_l
                                const float2 sc = make_float2(inv_s8, inv_s8);
                                const float2 lo2 = __ffma2_rn(make_float2(__uint_as_float(r[4 * q + 0]), __uint_as_float(r[4 * q + 1])), sc, make_float2(b4.x, b4.y));
                                const float2 hi2 = __ffma2_rn(make_float2(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])), sc, make_float2(b4.z, b4.w));
                                v[4 * q + 0] = lo2.x; v[4 * q + 1] = lo2.y; v[4 * q + 2] = hi2.x; v[4 * q + 3] = hi2.y;
                            } else {
                                const float2 lo2 = __fadd2_rn(make_float2(__uint_as_float(r[4 * q + 0]), __uint_as_float(r[4 * q + 1])), make_float2(b4.x, b4.y));
                                const float2 hi2 = __fadd2_rn(make_float2(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])), make_float2(b4.z, b4.w));
                                v[4 * q + 0] = lo2.x; v[4 * q + 1] = lo2.y; v[4 * q + 2] = hi2.x; v[4 * q + 3] = hi2.y;
                            }
                        }
                    } else {
                        const int n = c_n0 + 16 * g;
                        const float s1 = ((a.out_mode == OUT_NORMALISED) ? 1.f : nc.sd) * inv_s8;
                        const float2 sc = make_float2(s1, s1);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 s4 = *reinterpret_cast<const float4*>(s_s0 + n + 4 * q);
                            const float2 lo2 = __ffma2_rn(make_float2(__uint_as_float(r[4 * q + 0]), __uint_as_float(r[4 * q + 1])), sc, make_float2(s4.x, s4.y));
                            const float2 hi2 = __ffma2_rn(make_float2(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])), sc, make_float2(s4.z, s4.w));
                            v[4 * q + 0] = lo2.x; v[4 * q + 1] = lo2.y; v[4 * q + 2] = hi2.x; v[4 * q + 3] = hi2.y;
                        }
                    }
                    if (out_dst != DST_FINAL) {
                        if (do_relu) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = relu_nan(v[i]);
                        }
                        if (FMT != 0) {
#pragma unroll
                            for (int i = 0; i < 16; i += 2) vmax = fmax3_abs(vmax, v[i], v[i + 1]);
                        }
                        uint32_t w[16];  // [0..7] hi words, [8..15] second-tile words (see split16)
                        if (DBG & 16) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(v[i]);
                        } else {
                            split16<FMT>(v, w);
                        }
                        if (out_dst == DST_TMEM) {
                            tmem_st16(tbase + static_cast<uint32_t>(16 * g), w);  // in place: these 16 columns become the next layer's k-step
                        } else {
                            uint8_t* dst = sm + P.off_act + ((c_n0 >> 4) + g) * KSTEP_BYTES + row * 16;
                            *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                            *reinterpret_cast<uint4*>(dst + A_KG_BYTES) = make_uint4(w[4], w[5], w[6], w[7]);
                            *reinterpret_cast<uint4*>(dst + 2 * A_KG_BYTES) = make_uint4(w[8], w[9], w[10], w[11]);
                            *reinterpret_cast<uint4*>(dst + 3 * A_KG_BYTES) = make_uint4(w[12], w[13], w[14], w[15]);
                        }
                    } else if (OM == OM_CHI2) {
                        const int n = c_n0 + 16 * g;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float rr = (v[i] - s_obs[n + i]) * s_isig[n + i];  // isig = 0 on padding
                            chi = fmaf(rr, rr, chi);
                        }
                    } else if (OM == OM_ERROR) {
                        // emulator.py:185-191 fused: squared difference to this row's true signal and its amplitude, in the band
                        const int n = c_n0 + 16 * g;
                        const float* tr = a.truth + grow * static_cast<long long>(NO) + n;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float t = (grow < a.n && n + i < NO) ? __ldg(tr + i) : 0.f;
                            const float rr = (v[i] - t) * s_isig[n + i];
                            chi = fmaf(rr, rr, chi);
                            amp = fmaxf(amp, fabsf(t) * s_isig[n + i]);
                        }
                    } else if (!(DBG & 8)) {
                        const int col0 = 16 * g;  // column of v[0] within the chunk
                        float* op = a.out + grow * static_cast<long long>(NO) + c_n0;
                        if (staged && col0 >= shift && col0 + 16 <= shift + st_w) {  // inside the box: stage
#pragma unroll
                            for (int i = 0; i < 16; ++i) sstage[col0 + i] = v[i];
                        } else if (grow < a.n) {  // box boundary (or no box): stage what belongs to the box, store the rest
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int col = col0 + i;
                                if (staged && col >= shift && col < shift + st_w) sstage[col] = v[i];
                                else if (c_n0 + col < NO && !(SDBG & 2)) __stcs(op + col, v[i]);
                            }
                        }
                    }
                };
                if (g_lo < g_hi && !(DBG & 2)) {
#if VAE21_TC_EPI_SINGLE
                    // one instantiation of the group body (half the epilogue's code): the prefetched registers are moved, not renamed
                    uint32_t ra[16], rb[16];
                    tmem_ld16(tbase + static_cast<uint32_t>(16 * g_lo), ra);
#pragma unroll 1
                    for (int g = g_lo; g < g_hi; ++g) {
                        tmem_ld_wait();
                        if (g + 1 < g_hi) tmem_ld16(tbase + static_cast<uint32_t>(16 * (g + 1)), rb);
                        process(ra, g);
                        if (g + 1 < g_hi) {
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) ra[i] = rb[i];
                        }
                    }
#else
                    uint32_t ra[16], rb[16];
                    tmem_ld16(tbase + static_cast<uint32_t>(16 * g_lo), ra);
#pragma unroll 1
                    for (int g = g_lo; g < g_hi; g += 2) {
                        tmem_ld_wait();
                        if (g + 1 < g_hi) tmem_ld16(tbase + static_cast<uint32_t>(16 * (g + 1)), rb);
                        process(ra, g);
                        if (g + 1 < g_hi) {
                            tmem_ld_wait();
                            if (g + 2 < g_hi) tmem_ld16(tbase + static_cast<uint32_t>(16 * (g + 2)), ra);
                            process(rb, g + 1);
                        }
                    }
#endif
                }
                if (OM == OM_ROWS && out_dst == DST_FINAL) {
                    // the accumulator buffer is free as soon as it has been read: release it BEFORE the store hand-off below, which is
                    // then off the MMA issuers' critical path
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0 && c_qbuf >= 0) signal_mma(bar_q_empty(c_qbuf));
                }
                if (OM == OM_ROWS && out_dst == DST_FINAL && tma_launch && !(DBG & (2 | 8))) {
                    fence_async_smem();  // the staged words are read by the async proxy
                    // earlier chunks' boxes have been read: behind the barrier everybody may overwrite the OTHER staging buffer
                    if (issuer && !(SDBG & 4)) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
                    if (!(SDBG & 8)) asm volatile("bar.sync %0, 128;\n" ::"r"(4 + sub) : "memory");
                    if (issuer) {
                        const long long r4 = (tile * MT) / 4 + 8 * sub;  // first super-row of this sub-partition
                        const int sh = (4 - ((lane * NO) & 3)) & 3;     // this lane issues the box of residue `lane`
                        if (st_w > 0 && 4 * r4 < smaps.rows4 && !(SDBG & 1)) {
                            const uint32_t src = smem_u32(sm + P.off_stage + ((P.stage_bufs == 2) ? ((fin_cnt - 1u) & 1u) * (16 * P.stage_sq) : 0) +
                                                          (sub * 4 + lane) * P.stage_sq);
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(
                                             reinterpret_cast<uint64_t>(&smaps.m[C.st_map])),
                                         "r"(lane * NO + c_n0 + sh), "r"(static_cast<int>(r4)), "r"(src)
                                         : "memory");
                        }
                        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                    }
                }
                // publish: operand visible to the tensor pipe / accumulator buffer free
                if (out_dst == DST_SMEM) fence_async_smem();
                if (out_dst == DST_TMEM) tmem_st_wait();
                tc_fence_before();
                __syncwarp();  // every lane's writes / reads are done and fenced before the elected lane signals
                if (lane == 0) {
                    if (c_qbuf >= 0 && !(OM == OM_ROWS && out_dst == DST_FINAL)) signal_mma(bar_q_empty(c_qbuf));
                    if (out_dst != DST_FINAL) signal_mma(bar_act_ready(c_idx));
                }
            }
            if (OM == OM_CHI2 || OM == OM_ERROR) {
                // tile end: the column shares of a row combine their chi^2 partials
                float* chi_buf = s_chi + (tcount & 1u) * (EPS - 1) * 128;
                float* amp_buf = s_amp + (tcount & 1u) * (EPS - 1) * 128;
                if (half > 0) chi_buf[(half - 1) * 128 + row] = chi;
                if (OM == OM_ERROR && half > 0) amp_buf[(half - 1) * 128 + row] = amp;
                asm volatile("bar.sync 1, %0;\n" ::"n"(NEPI) : "memory");
                if (OM == OM_ERROR && half == 0) {
#pragma unroll
                    for (int hh = 0; hh < EPS - 1; ++hh) {
                        chi += chi_buf[hh * 128 + row];
                        amp = fmaxf(amp, amp_buf[hh * 128 + row]);
                    }
                    if (grow < a.n) {
                        float e = sqrtf(chi * a.err_inv_count);
                        if (a.err_relative) e = e / amp * 100.f;
                        a.chi2[grow] = e;
                    }
                }
                if (OM == OM_CHI2 && half == 0) {
#pragma unroll
                    for (int hh = 0; hh < EPS - 1; ++hh) chi += chi_buf[hh * 128 + row];
                    unsigned long long key = ~0ull;
                    if (grow < a.n) {
                        if (a.chi2) a.chi2[grow] = chi;
                        key = pack_min_key(chi, static_cast<unsigned long long>(a.row_base + grow));
                    }
                    if (a.argmin_key) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                            key = other < key ? other : key;
                        }
                        if (lane == 0 && key != ~0ull) atomicMin(a.argmin_key, key);
                    }
                }
            }
        }
        if (OM == OM_ROWS && issuer) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");  // all box stores complete before the CTA exits
        // operand-range check (FMT 1: fp16 hi/lo overflow beyond 65504; FMT 2: the e4m3 correction operands clip at 448): count the
        // epilogue threads that converted a hidden activation beyond the range.  fmaxf ignores NaN, so the reference's NaN
        // propagation for non-positive parameters does not count.
        if (FMT != 0 && a.sat) {
            const unsigned m = __ballot_sync(0xffffffffu, vmax > sat_limit<FMT>());
            if (lane == 0 && m) atomicAdd(a.sat, static_cast<unsigned long long>(__popc(m)));
        }
    }

    // ---- teardown -------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // no CTA frees TMEM / exits while its peer may still signal it
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512));
    }
}

template <int FMT, int CG, int OM>
inline cudaError_t launch_one(const Plan& P, const NormConsts& nc, const LaunchArgs& a, const uint8_t* wimg, const float* bias,
                              int grid, cudaStream_t st, const StoreMaps& smaps) {
    static unsigned long long prepared = 0;  // per instantiation: bit d = the shared-memory attribute is set on device d
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !((prepared >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(vae21_tc_kernel<FMT, CG, OM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) return e;
        if (dev < 64) prepared |= 1ull << dev;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = CG == 2 ? P.smem_total2 : P.smem_total;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, vae21_tc_kernel<FMT, CG, OM>, P, nc, a, wimg, bias, smaps);
}

inline cudaError_t prepare() { return cudaSuccess; }  // attributes are set per instantiation on first launch (launch_one)

// Tensor maps for the OM_ROWS store path of one launch (see StoreMaps).  Leaves rows4 = 0 -- plain stores -- when the output is not
// 16-byte aligned, has fewer than 4 rows, or the driver entry point is missing.  Maps are cached per (output pointer, rows, plan
// widths): the host pipeline re-uses three device buffers, benchmark loops one.
inline void make_store_maps(const Plan& P, const LaunchArgs& a, StoreMaps& sm) {
    std::memset(&sm, 0, sizeof(sm));
    if (a.out_mode == OUT_CHI2 || a.out_mode == OUT_ERROR || !a.out || P.n_maps == 0) return;
    if ((reinterpret_cast<uintptr_t>(a.out) & 15u) || a.n < 4) return;
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) fn = nullptr;
        return reinterpret_cast<encode_fn>(fn);
    }();
    if (!encode) return;
    struct Entry { const float* out; long long n; int n_out, n_maps, w[4]; StoreMaps sm; };
    static thread_local Entry cache[8];
    static thread_local int next = 0;
    for (const Entry& e : cache)
        if (e.out == a.out && e.n == a.n && e.n_out == P.n_out && e.n_maps == P.n_maps && std::memcmp(e.w, P.map_w, sizeof(e.w)) == 0) {
            sm = e.sm;
            return;
        }
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(4) * P.n_out, static_cast<cuuint64_t>(a.n / 4)};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(16) * P.n_out};
    const cuuint32_t estride[2] = {1, 1};
    for (int i = 0; i < P.n_maps; ++i) {
        const cuuint32_t box[2] = {static_cast<cuuint32_t>(P.map_w[i]), 8};
        if (encode(&sm.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.out, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            std::memset(&sm, 0, sizeof(sm));
            return;
        }
    }
    sm.rows4 = a.n / 4 * 4;
    Entry& e = cache[next];
    next = (next + 1) % 8;
    e.out = a.out; e.n = a.n; e.n_out = P.n_out; e.n_maps = P.n_maps;
    std::memcpy(e.w, P.map_w, sizeof(e.w));
    e.sm = sm;
}

template <int FMT, int CG>
inline cudaError_t launch_om(const Plan& P, const NormConsts& nc, const LaunchArgs& a, const uint8_t* w, const float* bias, int grid,
                             cudaStream_t st) {
    StoreMaps smaps;
    make_store_maps(P, a, smaps);
    if (a.out_mode == OUT_CHI2) return launch_one<FMT, CG, OM_CHI2>(P, nc, a, w, bias, grid, st, smaps);
    if (a.out_mode == OUT_ERROR) return launch_one<FMT, CG, OM_ERROR>(P, nc, a, w, bias, grid, st, smaps);
    return launch_one<FMT, CG, OM_ROWS>(P, nc, a, w, bias, grid, st, smaps);
}

// cta_group: 1 = one CTA per 128-row tile, 2 = CTA pairs on 256-row super-tiles
inline cudaError_t launch(const Plan& P, const NormConsts& nc, const LaunchArgs& a, const void* wimg, const float* bias,
                          int fmt, int sm_count, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
#ifdef VAE21_TC_DEV_FAST  // development builds: only the two pair kernels the A/B scripts run (compiles in a fraction of the time)
    {
        const long long nunits = (a.n + 2 * MT - 1) / (2 * MT);
        const int grid = 2 * static_cast<int>(std::min<long long>(nunits, sm_count / 2));
        if (a.out_mode == OUT_CHI2 || a.out_mode == OUT_ERROR || fmt == 1) return cudaErrorNotSupported;
        StoreMaps smaps;
        make_store_maps(P, a, smaps);
        return fmt == 0 ? launch_one<0, 2, OM_ROWS>(P, nc, a, static_cast<const uint8_t*>(wimg), bias, grid, st, smaps)
                        : launch_one<2, 2, OM_ROWS>(P, nc, a, static_cast<const uint8_t*>(wimg), bias, grid, st, smaps);
    }
#else
    static const int cg_env = std::getenv("VAE21_TC_CTA_GROUP") ? std::atoi(std::getenv("VAE21_TC_CTA_GROUP")) : 0;
    const int cg = (cg_env == 1 && P.nslots >= 2) ? 1 : P.default_cg;  // the one-CTA variant is an experiment switch and may not fit
    const uint8_t* w = static_cast<const uint8_t*>(wimg);
    if (cg == 2) {
        const long long nunits = (a.n + 2 * MT - 1) / (2 * MT);
        const int grid = 2 * static_cast<int>(std::min<long long>(nunits, sm_count / 2));
        return fmt == 0 ? launch_om<0, 2>(P, nc, a, w, bias, grid, st)
               : fmt == 1 ? launch_om<1, 2>(P, nc, a, w, bias, grid, st) : launch_om<2, 2>(P, nc, a, w, bias, grid, st);
    }
    const long long ntiles = (a.n + MT - 1) / MT;
    const int grid = static_cast<int>(std::min<long long>(ntiles, sm_count));
    return fmt == 0 ? launch_om<0, 1>(P, nc, a, w, bias, grid, st)
           : fmt == 1 ? launch_om<1, 1>(P, nc, a, w, bias, grid, st) : launch_om<2, 1>(P, nc, a, w, bias, grid, st);
#endif
}

#if VAE21_TC_TIMING
inline cudaError_t read_timing(long long* host /*[160*16]*/) { return cudaMemcpyFromSymbol(host, g_tc_timing, sizeof(long long) * 160 * 16); }
inline cudaError_t read_rec_timing(unsigned long long* host /*[3*256]*/, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(host, g_tc_rec, sizeof(unsigned long long) * 3 * 256);
    if (e == cudaSuccess && reset) {
        static unsigned long long zero[3 * 256] = {0};
        e = cudaMemcpyToSymbol(g_tc_rec, zero, sizeof(zero));
    }
    return e;
}
#endif

}  // namespace tck
