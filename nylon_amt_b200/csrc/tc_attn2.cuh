// Persistent, pipelined tcgen05 attention:  softmax(Q K^T / sqrt(dh)) V   for the hFT sequences (Lk <= 256, dh = 64)
// (reference MultiHeadAttentionLayer.forward, model_spec2midi.py:342-348).
//
// One CTA per SM walks "rounds" of two 128-query tiles.  Roles:
//   warp 0      TMA producer   Q tiles (one buffer per softmax warpgroup) and a ring of K / V half-tiles (128 keys x dh)
//   warp 1      MMA issuer     S = Q K^T into TMEM; O = P V with P read straight from TMEM (tcgen05.mma A-from-TMEM)
//   warps 2..5  softmax warpgroup 0, warps 6..9 softmax warpgroup 1: one thread per query row (TMEM lane), so the row
//               maximum / sum need no cross-thread exchange
// A tile's keys are processed in halves of 128 ("units").  Each unit has its own maximum and sum (split softmax):
//   S_u = Q K_u^T (128 TMEM columns) -> P_u = exp2(S_u c - m_u) written IN PLACE over S_u as 16-bit pairs
//   -> O_u = P_u V_u into its own 64 accumulator columns;  the epilogue merges  O = (O_a f_a + O_b f_b) / (l_a f_a + l_b f_b),
//   f_u = exp2(m_u - max(m_a, m_b)),
// so a warpgroup never holds more than 256 TMEM columns (S/P 128 | O_a 64 | O_b 64) and the two warpgroups ping-pong: while
// one waits for its MMAs the other runs exp2 on the CUDA cores.  With 256-query sequences (encoder) the two warpgroups
// take the two query tiles of the same (sequence, head) and SHARE every K / V half-tile in the ring, which halves the
// L2 -> shared-memory traffic; otherwise they work on different (sequence, head) items.
// x3 mode (split operands): S = Qh Kh + Ql Kh + Qh Kl;  P = Ph + Pl (both in TMEM, [hi 16 cols | lo 16 cols] per 32 keys);
// O = Ph Vh + Pl Vh + Ph Vl;  ctx stored hi | lo.
// PROBS variant (last decoder cross-attention, model_spec2midi.py:360): every unit's un-normalised exp2 values are written to the
// fp32 probability tensor during its softmax pass; once both units' maxima and sums are known the same thread rescales its own row
// in place (the row is still in L2), so no extra TMEM state is needed.
#pragma once
#include "tc_attn.cuh"

namespace hft {
namespace tc {

struct Attn2Params {
  int lq;                 // valid query rows per sequence
  int lk;                 // valid keys per sequence
  int q_seq_rows;         // rows between consecutive sequences in the Q tensor (0 = the same queries for every sequence)
  int q_tiles;            // ceil(lq / 128): 1 or 2
  int heads;
  int q_col0, k_col0, v_col0;
  float scale_log2e;
  void* ctx;              // 16-bit [n_seq * lq, ld_ctx]
  int ld_ctx;
  int q_lo_off, kv_lo_off, ctx_lo_off;
  int tma_store;          // 1: ctx written by per-warp TMA stores (lq % 128 == 0)
  int n_items;            // n_seq * heads * q_tiles
  int n_rounds;           // ceil(n_items / 2)
  int shared_kv;          // 1: q_tiles == 2, both warpgroups use the same K / V half-tiles
  float* probs;           // PROBS: fp32 [n_seq, heads, lq, lk]
  int s_single;           // X3 layout, mixed-precision plan: S = Qh Kh only (the lo halves of Q and K are not loaded)
  int pv_single;          // X3 layout, mixed-precision plan: O = Ph Vh only (the lo half of V is not loaded; P is still written hi | lo)
};

template <int DH, bool X3>
struct Attn2Smem {
  static constexpr int parts = X3 ? 2 : 1;
  static constexpr int tile_bytes = 128 * DH * 2;                 // one 128-row x dh operand part
  static constexpr int q_bytes = tile_bytes * parts;              // per warpgroup
  static constexpr int slot_bytes = tile_bytes * parts;           // one K or V half-tile (hi | lo)
  static constexpr int n_slots = X3 ? 4 : 8;
  static constexpr int stage_bytes = 8 * 32 * DH * 2;             // per warp one 32-row block (hi, then lo)
  static constexpr int total = 1024 + 2 * q_bytes + n_slots * slot_bytes + stage_bytes + 512;
};

constexpr int kAttn2Threads = 64 + 256;

// NKEY: keys per unit (128, or 96 for the 88-key decoder self-attention); NH: units per tile (Lk = NH * NKEY)
template <bool BF16, int DH, int NKEY, int NH, bool X3, bool PROBS = false>
__global__ void __launch_bounds__(kAttn2Threads, 1) attn2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                                                                const __grid_constant__ CUtensorMap map_o, const __grid_constant__ Attn2Params p) {
  using L = Attn2Smem<DH, X3>;
  static_assert(DH == 64, "attn2 is built for head_dim 64");
  static_assert((NKEY == 128 || NKEY == 96) && (NH == 1 || NH == 2), "unsupported key tiling");
  constexpr int kParts = X3 ? 2 : 1;
  constexpr int NS = L::n_slots;
  constexpr uint32_t kRowBytes = DH * 2;                  // 128
  constexpr uint32_t kAtom = 8 * kRowBytes;               // 1024
  constexpr uint32_t kOCol = 128;                         // O_a at +128, O_b at +192 inside a warpgroup's 256 columns
  constexpr int kChunks = NKEY / 32;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_q = smem;                                                   // [2 warpgroups][parts][128 x dh]
  uint8_t* s_ring = s_q + 2 * L::q_bytes;                                // [NS][parts][128 x dh]
  uint8_t* s_stage = s_ring + NS * L::slot_bytes;                        // [8 warps][32 x dh]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + L::stage_bytes);
  uint64_t* ring_full = bars;            // [8]
  uint64_t* ring_empty = bars + 8;       // [8]
  uint64_t* q_full = bars + 16;          // [2]
  uint64_t* q_empty = bars + 18;         // [2]
  uint64_t* s_ready = bars + 20;         // [2]  S unit complete (MMA -> softmax)
  uint64_t* p_ready = bars + 22;         // [2]  P unit written (softmax -> MMA), 4 warp arrivals
  uint64_t* o_ready = bars + 24;         // [2]  all PV MMAs of the tile complete (MMA -> epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_kv); tma_prefetch_desc(&map_o);
    for (int i = 0; i < NS; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&s_ready[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&o_ready[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const bool shared = p.shared_kv != 0;

  // tile index -> (sequence, head, query tile)
  auto decode = [&](int tile, int& seq, int& head, int& qt) {
    qt = tile % p.q_tiles;
    head = (tile / p.q_tiles) % p.heads;
    seq = tile / (p.q_tiles * p.heads);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0, qph[2] = {0, 0};
      const int k_parts = (X3 && !p.s_single) ? 2 : 1, v_parts = (X3 && !p.pv_single) ? 2 : 1;   // operand halves actually multiplied
      auto load_kv = [&](int tile, int col0, int half) {                  // one K or V half-tile into the next ring slot
        int seq, head, qt;
        decode(tile, seq, head, qt);
        const int np = col0 == p.v_col0 ? v_parts : k_parts;
        mbar_wait(&ring_empty[slot], ph ^ 1);
        mbar_expect_tx(&ring_full[slot], (uint32_t)(NKEY * DH * 2 * np));
        for (int part = 0; part < np; ++part)
          tma_load_2d(s_ring + (size_t)slot * L::slot_bytes + part * L::tile_bytes, &map_kv, part * p.kv_lo_off + col0 + head * DH, seq * p.lk + half * NKEY,
                      &ring_full[slot]);
        if (++slot == NS) { slot = 0; ph ^= 1; }
      };
      auto load_q = [&](int w, int tile) {
        int seq, head, qt;
        decode(tile, seq, head, qt);
        mbar_wait(&q_empty[w], qph[w] ^ 1);
        mbar_expect_tx(&q_full[w], (uint32_t)(L::tile_bytes * k_parts));
        for (int part = 0; part < k_parts; ++part)
          tma_load_2d(s_q + (size_t)w * L::q_bytes + part * L::tile_bytes, &map_q, part * p.q_lo_off + p.q_col0 + head * DH, seq * p.q_seq_rows + qt * 128, &q_full[w]);
        qph[w] ^= 1;
      };
      int prev[2] = {-1, -1};
      for (int round = blockIdx.x; round < p.n_rounds; round += gridDim.x) {
        int tile[2] = {2 * round, 2 * round + 1 < p.n_items ? 2 * round + 1 : -1};
        for (int w = 0; w < 2; ++w) {                                     // phase A: last V half of the previous tile, first K half of the new one
          if (prev[w] >= 0 && (!shared || w == 0)) load_kv(prev[w], p.v_col0, NH - 1);
          if (tile[w] >= 0) {
            load_q(w, tile[w]);
            if (!shared || w == 0) load_kv(tile[w], p.k_col0, 0);
          }
        }
        if (NH == 2) {
          for (int w = 0; w < 2; ++w)                                     // phase B: first V half, second K half
            if (tile[w] >= 0 && (!shared || w == 0)) { load_kv(tile[w], p.v_col0, 0); load_kv(tile[w], p.k_col0, 1); }
        }
        prev[0] = tile[0]; prev[1] = tile[1];
      }
      for (int w = 0; w < 2; ++w)
        if (prev[w] >= 0 && (!shared || w == 0)) load_kv(prev[w], p.v_col0, NH - 1);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {                                                       // the whole warp runs the loop (uniform control flow); one elected lane issues MMAs and commits
      const uint32_t idesc_s = make_idesc(128, NKEY, BF16, false, false);   // S = Q K^T, both K-major
      const uint32_t idesc_o = make_idesc(128, DH, BF16, false, true);      // O = P V, A from TMEM, B = V MN-major
      const SDescBase kds = sdesc_base(16, kAtom, kSwz128), kdv = sdesc_base(kAtom, kAtom, kSwz128);   // one add per descriptor in the loops below
      auto commit = [&](uint64_t* bar) { if (elect_one()) umma_commit(bar); __syncwarp(); };
      int slot = 0;
      uint32_t ph = 0, qph[2] = {0, 0}, pph[2] = {0, 0};
      auto take = [&]() -> int {
        mbar_wait(&ring_full[slot], ph);
        fence_after_sync();
        const int s = slot;
        if (++slot == NS) { slot = 0; ph ^= 1; }
        return s;
      };
      auto issue_s = [&](int w, int kslot) {                                // S(w)[128, NKEY] = Q(w) K^T;  x3: Qh Kh + Ql Kh + Qh Kl
        const uint32_t qa = smem_u32(s_q + (size_t)w * L::q_bytes), ka = smem_u32(s_ring + (size_t)kslot * L::slot_bytes);
        const int n_prod = (X3 && !p.s_single) ? 3 : 1;
        if (elect_one()) {
          uint32_t acc = 0;
          for (int part = 0; part < n_prod; ++part) {
            const uint32_t qp = qa + (part == 1 ? L::tile_bytes : 0), kp = ka + (part == 2 ? L::tile_bytes : 0);
            const uint32_t q0 = sdesc_lo(kds, qp), k0 = sdesc_lo(kds, kp);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) {
              umma_f16_lohi(tmem_base + w * 256, q0 + 2 * k, kds.hi, k0 + 2 * k, kds.hi, idesc_s, acc);
              acc = 1;
            }
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](int w, uint32_t ocol, int vslot) {                // O_u(w)[128, dh] = P V;  x3: Ph Vh + Pl Vh + Ph Vl
        const uint32_t va = smem_u32(s_ring + (size_t)vslot * L::slot_bytes);
        const uint32_t tp = tmem_base + w * 256;
        const int n_prod = (X3 && !p.pv_single) ? 3 : 1;
        if (elect_one()) {
          uint32_t acc = 0;
          for (int part = 0; part < n_prod; ++part) {
            const uint32_t v0 = sdesc_lo(kdv, va + (part == 2 ? L::tile_bytes : 0));
#pragma unroll
            for (int k = 0; k < NKEY / 16; ++k) {
              // P columns: x3 [hi 16 | lo 16] per 32 keys; single product 8 columns per 16 keys
              const uint32_t pcol = X3 ? (uint32_t)((k >> 1) * 32 + (part == 1 ? 16 : 0) + (k & 1) * 8) : (uint32_t)(k * 8);
              umma_f16_ts_lohi(tp + ocol, tp + pcol, v0 + k * (16 * kRowBytes / 16), kdv.hi, idesc_o, acc);
              acc = 1;
            }
          }
        }
        __syncwarp();
      };
      int prev[2] = {-1, -1};
      int held_v = 0, held_k = 0;
      auto finish_prev = [&](int w) {                                        // PV of the previous tile's last unit
        if (!shared || w == 0) held_v = take();
        mbar_wait(&p_ready[w], pph[w]); pph[w] ^= 1;
        fence_after_sync();
        issue_pv(w, NH == 2 ? kOCol + 64 : kOCol, held_v);
        commit(&o_ready[w]);
        if (!shared || w == 1) commit(&ring_empty[held_v]);
      };
      for (int round = blockIdx.x; round < p.n_rounds; round += gridDim.x) {
        int tile[2] = {2 * round, 2 * round + 1 < p.n_items ? 2 * round + 1 : -1};
        for (int w = 0; w < 2; ++w) {
          if (prev[w] >= 0) finish_prev(w);
          if (tile[w] >= 0) {
            if (!shared || w == 0) held_k = take();
            mbar_wait(&q_full[w], qph[w]); qph[w] ^= 1;
            fence_after_sync();
            issue_s(w, held_k);
            commit(&s_ready[w]);
            if (NH == 1) commit(&q_empty[w]);
            if (!shared || w == 1) commit(&ring_empty[held_k]);
          }
        }
        if (NH == 2) {
          for (int w = 0; w < 2; ++w) {
            if (tile[w] < 0) continue;
            if (!shared || w == 0) { held_v = take(); held_k = take(); }
            mbar_wait(&p_ready[w], pph[w]); pph[w] ^= 1;
            fence_after_sync();
            issue_pv(w, kOCol, held_v);
            issue_s(w, held_k);                                              // overwrites P_a: the MMA pipe runs in issue order
            commit(&s_ready[w]);
            commit(&q_empty[w]);
            if (!shared || w == 1) { commit(&ring_empty[held_v]); commit(&ring_empty[held_k]); }
          }
        }
        prev[0] = tile[0]; prev[1] = tile[1];
      }
      for (int w = 0; w < 2; ++w)
        if (prev[w] >= 0) finish_prev(w);
    }
  } else {
    // ===================== softmax + epilogue warpgroups =====================
    const int w = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                       // query row inside the tile == TMEM lane
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + w * 256;
    uint8_t* my_stage = s_stage + (size_t)(warp - 2) * (32 * DH * 2);
    uint32_t sph = 0, oph = 0;
    for (int round = blockIdx.x; round < p.n_rounds; round += gridDim.x) {
      const int tile = 2 * round + w;
      if (tile >= p.n_items) continue;
      int seq, head, qt;
      decode(tile, seq, head, qt);
      const int qrow_p = qt * 128 + r;
      float* prow = PROBS ? p.probs + (((long long)seq * p.heads + head) * p.lq + (qrow_p < p.lq ? qrow_p : 0)) * p.lk : nullptr;
      float mx[NH], sum[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        mbar_wait(&s_ready[w], sph); sph ^= 1;
        fence_after_sync();
        const int kvalid = p.lk - h * NKEY;                  // valid keys in this unit (only the 88-key sequences are padded)
        const bool mask = kvalid < NKEY;
        float m = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          uint32_t v[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
          if (mask) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < kvalid) m = fmaxf(m, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
          }
        }
        const float ms = m * p.scale_log2e;
        float s = 0.f;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          uint32_t v[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
          uint32_t pw[32];                                   // x3: [hi 16 | lo 16]; single product: first 16 used
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), p.scale_log2e, -ms));
            float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2e, -ms));
            if (mask) {
              if (c * 32 + 2 * j >= kvalid) e0 = 0.f;
              if (c * 32 + 2 * j + 1 >= kvalid) e1 = 0.f;
            }
            s += e0 + e1;
            if (PROBS) { v[2 * j] = __float_as_uint(e0); v[2 * j + 1] = __float_as_uint(e1); }
            if (X3) split_pack<BF16>(e0, e1, pw[j], pw[16 + j]);
            else pw[j] = Op16<BF16>::pack(e0, e1);
          }
          if (X3) {
            tmem_st32(t_row + c * 32, pw);
          } else {
            uint32_t ph16[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) ph16[j] = pw[j];
            tmem_st16(t_row + c * 16, ph16);
          }
          if (PROBS && qrow_p < p.lq) {                     // un-normalised for now; rescaled below once max / sum of the whole row are known
            float4* dst = reinterpret_cast<float4*>(prow + h * NKEY + c * 32);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4)
              dst[q4] = make_float4(__uint_as_float(v[4 * q4]), __uint_as_float(v[4 * q4 + 1]), __uint_as_float(v[4 * q4 + 2]), __uint_as_float(v[4 * q4 + 3]));
          }
        }
        mx[h] = ms;
        sum[h] = s;
        tmem_st_wait();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[w]);
      }
      // ---- epilogue: merge the units, normalise, store ----
      float f0 = 1.f, f1 = 0.f;
      if (NH == 2) {
        const float mm = fmaxf(mx[0], mx[NH - 1]);
        f0 = ex2_approx(mx[0] - mm);
        f1 = ex2_approx(mx[NH - 1] - mm);
      }
      const float inv = 1.f / (NH == 2 ? sum[0] * f0 + sum[NH - 1] * f1 : sum[0]);
      f0 *= inv; f1 *= inv;
      if (PROBS && qrow_p < p.lq) {                         // softmax probabilities: this thread's own row, written above
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const float f = h == 0 ? f0 : f1;
          float4* row4 = reinterpret_cast<float4*>(prow + h * NKEY);
#pragma unroll 4
          for (int q4 = 0; q4 < NKEY / 4; ++q4) {
            float4 t = row4[q4];
            row4[q4] = make_float4(t.x * f, t.y * f, t.z * f, t.w * f);
          }
        }
      }
      mbar_wait(&o_ready[w], oph); oph ^= 1;
      fence_after_sync();
      const int qrow = qt * 128 + r;
      uint32_t hi[32], lo[32];                               // the row's dh = 64 outputs as 16-bit pairs
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) {
        uint32_t oa[32], ob[32];
        tmem_ld32(t_row + kOCol + c * 32, oa);
        if (NH == 2) tmem_ld32(t_row + kOCol + 64 + c * 32, ob);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = __uint_as_float(oa[2 * j]) * f0, b = __uint_as_float(oa[2 * j + 1]) * f0;
          if (NH == 2) { a = fmaf(__uint_as_float(ob[2 * j]), f1, a); b = fmaf(__uint_as_float(ob[2 * j + 1]), f1, b); }
          if (X3) split_pack<BF16>(a, b, hi[c * 16 + j], lo[c * 16 + j]);
          else hi[c * 16 + j] = Op16<BF16>::pack(a, b);
        }
      }
      if (p.tma_store) {
        const int row0 = seq * p.lq + qt * 128 + quarter * 32;
#pragma unroll
        for (int part = 0; part < kParts; ++part) {
          if (lane == 0) tma_store_wait_read();              // the warp's previous store has read the staging block
          __syncwarp();
          uint8_t* dst = my_stage + lane * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(dst + ((q ^ (lane & 7)) << 4)) =
                part ? make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]) : make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_o, my_stage, part * p.ctx_lo_off + head * DH, row0);
            tma_store_commit();
          }
        }
      } else if (qrow < p.lq) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.ctx) + ((long long)seq * p.lq + qrow) * p.ld_ctx + head * DH);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          *reinterpret_cast<uint4*>(dst + q * 4) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
          if (X3) *reinterpret_cast<uint4*>(dst + p.ctx_lo_off / 2 + q * 4) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Probabilities-returning attention (last decoder cross-attention, model_spec2midi.py:164-165, :360): Lq <= 128 queries against 256 keys.
// The returned softmax needs the row maximum / sum over ALL keys before anything is written, so this variant keeps the whole 256-column
// score row of a tile in TMEM (S/P 256 columns + O 64) and runs ONE softmax warpgroup; it is still persistent and pipelined (Q and the
// K / V tiles of the next item are loaded behind the current item's softmax; the PV MMA overlaps the probability write-out, which
// re-reads P = hi + lo from TMEM once the row sum is known).  Warps: 0 producer, 1 MMA, 2..5 softmax / epilogue (thread = query row).
// ---------------------------------------------------------------------------------------------------------------------------------
template <bool X3>
struct AttnProbsSmem {
  static constexpr int parts = X3 ? 2 : 1;
  static constexpr int q_bytes = 128 * 64 * 2 * parts;
  static constexpr int kv_part = 256 * 64 * 2;                    // one 256-key x 64 operand part
  static constexpr int slot_bytes = kv_part * parts;
  static constexpr int total = 1024 + q_bytes + 2 * slot_bytes + 256;
};
constexpr int kAttnProbsThreads = 64 + 128;

template <bool BF16, bool X3>
__global__ void __launch_bounds__(kAttnProbsThreads, 1) attn_probs_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                                                                         const __grid_constant__ Attn2Params p) {
  using L = AttnProbsSmem<X3>;
  constexpr int kParts = X3 ? 2 : 1;
  constexpr int DH = 64, LK = 256;
  constexpr uint32_t kAtom = 1024, kRowBytes = 128, kOCol = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_q = smem;
  uint8_t* s_ring = s_q + L::q_bytes;                                   // [2][parts][256 x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + 2 * L::slot_bytes);
  uint64_t* ring_full = bars;          // [2]
  uint64_t* ring_empty = bars + 2;     // [2]
  uint64_t* q_full = bars + 4;
  uint64_t* q_empty = bars + 5;
  uint64_t* s_ready = bars + 6;
  uint64_t* p_ready = bars + 7;        // 4 warp arrivals
  uint64_t* p_done = bars + 8;         // 4 warp arrivals: the probability write-out has finished reading P
  uint64_t* o_ready = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_kv);
    for (int i = 0; i < 2; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
    mbar_init(q_full, 1); mbar_init(q_empty, 1); mbar_init(s_ready, 1); mbar_init(p_ready, 4); mbar_init(p_done, 4); mbar_init(o_ready, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0, qph = 0;
      for (int tile = blockIdx.x; tile < p.n_items; tile += gridDim.x) {
        const int head = tile % p.heads, seq = tile / p.heads;
        const int k_parts = (X3 && !p.s_single) ? 2 : 1, v_parts = (X3 && !p.pv_single) ? 2 : 1;
        mbar_wait(q_empty, qph ^ 1);
        mbar_expect_tx(q_full, (uint32_t)(128 * 64 * 2 * k_parts));
        for (int part = 0; part < k_parts; ++part)
          tma_load_2d(s_q + part * (128 * 64 * 2), &map_q, part * p.q_lo_off + p.q_col0 + head * DH, seq * p.q_seq_rows, q_full);
        qph ^= 1;
        for (int kv = 0; kv < 2; ++kv) {                                  // K, then V
          const int np = kv ? v_parts : k_parts;
          mbar_wait(&ring_empty[slot], ph ^ 1);
          mbar_expect_tx(&ring_full[slot], (uint32_t)(L::kv_part * np));
          for (int part = 0; part < np; ++part)
            tma_load_2d(s_ring + (size_t)slot * L::slot_bytes + part * L::kv_part, &map_kv, part * p.kv_lo_off + (kv ? p.v_col0 : p.k_col0) + head * DH, seq * p.lk,
                        &ring_full[slot]);
          if (++slot == 2) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {                                                       // whole warp, uniform control flow; one elected lane issues MMAs and commits
      const uint32_t idesc_s = make_idesc(128, LK, BF16, false, false);
      const uint32_t idesc_o = make_idesc(128, DH, BF16, false, true);
      const SDescBase kds2 = sdesc_base(16, kAtom, kSwz128), kdv2 = sdesc_base(kAtom, kAtom, kSwz128);
      auto commit = [&](uint64_t* bar) { if (elect_one()) umma_commit(bar); __syncwarp(); };
      int slot = 0;
      uint32_t ph = 0, qph = 0, pph = 0, dph = 0;
      auto take = [&]() -> int {
        mbar_wait(&ring_full[slot], ph);
        fence_after_sync();
        const int s = slot;
        if (++slot == 2) { slot = 0; ph ^= 1; }
        return s;
      };
      for (int tile = blockIdx.x; tile < p.n_items; tile += gridDim.x) {
        const int ks = take();
        mbar_wait(q_full, qph); qph ^= 1;
        mbar_wait(p_done, dph ^ 1); dph ^= 1;                             // the previous item's probabilities have been read out of P
        fence_after_sync();
        const uint32_t qa = smem_u32(s_q), ka = smem_u32(s_ring + (size_t)ks * L::slot_bytes);
        const int s_prod = (X3 && !p.s_single) ? 3 : 1, pv_prod = (X3 && !p.pv_single) ? 3 : 1;
        if (elect_one()) {
          uint32_t acc = 0;
          for (int part = 0; part < s_prod; ++part) {                     // S = Qh Kh + Ql Kh + Qh Kl
            const uint32_t qp = qa + (part == 1 ? 128 * 64 * 2 : 0), kp = ka + (part == 2 ? L::kv_part : 0);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) {
              umma_f16_lohi(tmem_base, sdesc_lo(kds2, qp) + 2 * k, kds2.hi, sdesc_lo(kds2, kp) + 2 * k, kds2.hi, idesc_s, acc);
              acc = 1;
            }
          }
        }
        __syncwarp();
        commit(s_ready);
        commit(q_empty);
        commit(&ring_empty[ks]);
        const int vs = take();
        mbar_wait(p_ready, pph); pph ^= 1;
        fence_after_sync();
        const uint32_t va = smem_u32(s_ring + (size_t)vs * L::slot_bytes);
        if (elect_one()) {
          uint32_t acc = 0;
          for (int part = 0; part < pv_prod; ++part) {                    // O = Ph Vh + Pl Vh + Ph Vl, A = P from TMEM
            const uint32_t vp = va + (part == 2 ? L::kv_part : 0);
#pragma unroll
            for (int k = 0; k < LK / 16; ++k) {
              const uint32_t pcol = X3 ? (uint32_t)((k >> 1) * 32 + (part == 1 ? 16 : 0) + (k & 1) * 8) : (uint32_t)(k * 8);
              umma_f16_ts_lohi(tmem_base + kOCol, tmem_base + pcol, sdesc_lo(kdv2, vp) + k * (16 * kRowBytes / 16), kdv2.hi, idesc_o, acc);
              acc = 1;
            }
          }
        }
        __syncwarp();
        commit(o_ready);
        commit(&ring_empty[vs]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t sph = 0, oph = 0;
    for (int tile = blockIdx.x; tile < p.n_items; tile += gridDim.x) {
      const int head = tile % p.heads, seq = tile / p.heads;
      const bool live = r < p.lq;
      mbar_wait(s_ready, sph); sph ^= 1;
      fence_after_sync();
      float m = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < LK / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(t_row + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
      }
      const float ms = m * p.scale_log2e;
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < LK / 32; ++c) {
        uint32_t v[32], pw[32];
        tmem_ld32(t_row + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), p.scale_log2e, -ms));
          const float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2e, -ms));
          sum += e0 + e1;
          if (X3) split_pack<BF16>(e0, e1, pw[j], pw[16 + j]);
          else pw[j] = Op16<BF16>::pack(e0, e1);
        }
        if (X3) {
          tmem_st32(t_row + c * 32, pw);
        } else {
          uint32_t ph16[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) ph16[j] = pw[j];
          tmem_st16(t_row + c * 16, ph16);
        }
      }
      tmem_st_wait();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      // probabilities: P = hi (+ lo) read back from TMEM, normalised, fp32 [n_seq, heads, lq, lk]
      const float inv = 1.f / sum;
      float* prow = p.probs + (((long long)seq * p.heads + head) * p.lq + (live ? r : 0)) * p.lk;
#pragma unroll 1
      for (int c = 0; c < LK / 32; ++c) {
        float e[32];
        if (X3) {
          uint32_t v[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            e[2 * j] = Op16<BF16>::lo(v[j]) + Op16<BF16>::lo(v[16 + j]);
            e[2 * j + 1] = Op16<BF16>::hi(v[j]) + Op16<BF16>::hi(v[16 + j]);
          }
        } else {
          uint32_t v[32];
          tmem_ld32(t_row + (c >> 1) * 32, v);                           // 16 packed columns per 32 keys: chunk c sits in half (c & 1) of this load
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t w = (c & 1) ? v[16 + j] : v[j];
            e[2 * j] = Op16<BF16>::lo(w);
            e[2 * j + 1] = Op16<BF16>::hi(w);
          }
        }
        if (live) {
          float4* dst = reinterpret_cast<float4*>(prow + c * 32);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) dst[q4] = make_float4(e[4 * q4] * inv, e[4 * q4 + 1] * inv, e[4 * q4 + 2] * inv, e[4 * q4 + 3] * inv);
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_done);
      // context
      mbar_wait(o_ready, oph); oph ^= 1;
      fence_after_sync();
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) {
        uint32_t o[32];
        tmem_ld32(t_row + kOCol + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = __uint_as_float(o[2 * j]) * inv, b = __uint_as_float(o[2 * j + 1]) * inv;
          if (X3) split_pack<BF16>(a, b, hi[c * 16 + j], lo[c * 16 + j]);
          else hi[c * 16 + j] = Op16<BF16>::pack(a, b);
        }
      }
      if (live) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.ctx) + ((long long)seq * p.lq + r) * p.ld_ctx + head * DH);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          *reinterpret_cast<uint4*>(dst + q * 4) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
          if (X3) *reinterpret_cast<uint4*>(dst + p.ctx_lo_off / 2 + q * 4) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace hft
