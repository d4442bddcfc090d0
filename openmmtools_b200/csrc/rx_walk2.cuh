// rx_walk2.cuh -- second-generation swap-all walker for power-of-two K <= 256 (included by rx_mix.cu).
//
// Same algorithm and the same results as k_mix_walk_pow2<U_FILTER24> (replicaexchange.py:321-349 bit for bit); what
// changes is the length of the dependent chain of one speculation round:
//   * lane <-> slot mapping is FIXED (lane = slot mod 32) and the window [h, h+32) rotates over the lanes, so a lane
//     keeps its slot record in registers until the slot leaves the window and the record of its next slot (s + 32)
//     is prefetched a whole window ahead: no shared-memory ring access on the chain (v1: LDS ring -> LDS perm -> LDS
//     image; here: LDS perm -> LDS image).  Ballots are rotated into window order with one funnel shift;
//   * records are 16 bytes {i|j<<16, back-mask, f32 log-uniform of the NEXT slot, alternatives}; the f64 log-uniform the
//     exact path needs is recomputed from the two words of the stream (same log() as the pre-pass, same value);
//   * the commit's bookkeeping (log entry, counters, promotion of the lanes that left the window) is issued after the
//     next round's permutation loads, in their latency shadow;
//   * there is no producer warp and no shared-memory ring (v1's was synchronised by volatile flags only): a lane reads the
//     records of its own slot class straight from global memory, two windows ahead (L2 latency is covered by more than
//     one round), and prefetches the line it will need 1024 slots later into L2 (DRAM latency).
// The kernel runs the bulk of a pass; the last < 161 slots / < 33 attempts of a pass are left to
// k_mix_walk_pow2<U_FILTER24, true>, which reads the same records.
#pragma once
#ifndef W2_ZDEP
#define W2_ZDEP 0
#endif

struct SlotRec2 {        // 16 bytes, one 2-word slot of the stream, state independent
    uint32_t ij;         // i | j << 16
    uint32_t backmask;   // bit 31-b: slot s-1-b shares a replica index with slot s (b = 0..30)
    uint32_t lu_next;    // float bits: log of the uniform the two words of slot s+1 would produce, rounded to f32
    uint32_t alts;       // a1_i | a1_j << 8: partner of i (of j) in the most recent earlier slot that touches it (else itself)
};

#define W2_LOOKBACK 48

// f64 log-uniform of slot s, exactly as k_slots_build / k_slots_build2 compute it
__device__ __forceinline__ double slot_logU(const uint32_t *__restrict__ words, unsigned s) {
    const double U = mt_double(words[2 * (size_t)s], words[2 * (size_t)s + 1]);
    return (U == 0.0) ? LOGU_ZERO : log(U);
}

__global__ void __launch_bounds__(256) k_slots_build2(const uint32_t *__restrict__ words, long long nslots, uint32_t mask,
                                                      SlotRec2 *__restrict__ rec) {
    __shared__ uint32_t s_ij[256 + W2_LOOKBACK];
    const long long s0 = (long long)blockIdx.x * 256;
    const int t = threadIdx.x;
    for (int q = t; q < 256 + W2_LOOKBACK; q += 256) {   // tile: slots s0-48 .. s0+255
        const long long s = s0 - W2_LOOKBACK + q;
        uint32_t ij = 0xffffffffu;
        if (s >= 0 && s < nslots) ij = (words[2 * s] & mask) | ((words[2 * s + 1] & mask) << 16);
        s_ij[q] = ij;
    }
    __syncthreads();
    const long long s = s0 + t;
    if (s >= nslots) return;
    const uint32_t ij = s_ij[t + W2_LOOKBACK];
    const uint32_t i = ij & 0xffffu, j = ij >> 16;
    uint32_t bm = 0, a1i = i, a1j = j;
    bool fi = false, fj = false;
#pragma unroll 4
    for (int b = 0; b < W2_LOOKBACK - 1; b++) {
        const uint32_t o = s_ij[t + W2_LOOKBACK - 1 - b];
        const uint32_t oi = o & 0xffffu, oj = o >> 16;
        const bool valid = (o != 0xffffffffu);
        if (b < 31) {
            const bool hit = valid && (oi == i || oi == j || oj == i || oj == j);
            bm |= (hit ? 1u : 0u) << (31 - b);
        }
        if (valid && oi != oj) {   // a slot with i == j never changes the permutation
            if (!fi && (oi == i || oj == i)) { a1i = (oi == i) ? oj : oi; fi = true; }
            if (!fj && (oi == j || oj == j)) { a1j = (oi == j) ? oj : oi; fj = true; }
        }
    }
    SlotRec2 r;
    r.ij = ij;
    r.backmask = bm;
    float lu = 0.f;
    if (s + 1 < nslots) lu = (float)slot_logU(words, (unsigned)(s + 1));
    r.lu_next = __float_as_uint(lu);
    r.alts = (a1i & 0xffu) | ((a1j & 0xffu) << 8);
    rec[s] = r;
}

// The filter's decision for one (i, j, si, sj): image values f_xy = image of u[x, s_y] - rowmin_x.  Same arithmetic as
// k_mix_walk_pow2<U_FILTER24> (see the bound's derivation there); eps0 = rowabs[i] + rowabs[j] + 1e-9.
__device__ __forceinline__ void w2_filter(float f_ii, float f_ij, float f_jj, float f_ji, float eps0, float lu, bool i_eq_j,
                                          bool &ge0, bool &acc, bool &undecided) {
    const float lp = (f_ii - f_ij) + (f_jj - f_ji);
    const float mag = (fabsf(f_ii) + fabsf(f_ij)) + (fabsf(f_jj) + fabsf(f_ji));
    const float eps = fmaf(mag, 3.2e-5f, eps0);
    const float d = lp - lu;
    const float mar = fmaf(fabsf(lp) + fabsf(lu), 1.3e-7f, eps);
    const bool dec_lp = fabsf(lp) > eps, dec_d = fabsf(d) > mar;
    const bool same = i_eq_j && (fabsf(f_ii) <= 3.0e38f);
    ge0 = (dec_lp && lp > 0.f) || same;
    acc = ge0 || (dec_lp && dec_d && d > 0.f);
    undecided = !(ge0 || (dec_lp && dec_d));
}

// One entry per replica k: the walker's view of the permutation.  `diag` is the image value of u[k, state] - rowmin_k, so
// a round reads the two diagonal terms of log_p together with the states (one dependent shared-memory level less) and
// only the two off-diagonal image values afterwards; `rowabs` is the row's share of the filter's rounding bound.
struct __align__(16) W2Replica {
    int state;
    float diag;
    float rowabs;
    int pad;
};

__device__ __forceinline__ float w2_image(const unsigned short *__restrict__ s_qhi, const unsigned char *__restrict__ s_qlo, unsigned a) {
    return __uint_as_float(__byte_perm((unsigned)s_qhi[a], (unsigned)s_qlo[a], 0x1045));   // (hi << 16) | (lo << 8)
}

__global__ void __launch_bounds__(32) k_mix_walk2(const SlotRec2 *__restrict__ rec, const uint32_t *__restrict__ words,
                                                  unsigned nslots, const double *__restrict__ u, int K, int logK,
                                                  int *__restrict__ perm_g, uint32_t *__restrict__ slot_log,
                                                  const unsigned char *__restrict__ filt,
                                                  const double *__restrict__ filt_rowabs, MixCtl *ctl) {
    extern __shared__ uint4 s_w2[];
    W2Replica *s_rep = (W2Replica *)s_w2;                  // [K]
    unsigned char *s_q = (unsigned char *)(s_rep + K);     // image: u16 plane [K*K], then u8 plane [K*K]
    const unsigned short *s_qhi = (const unsigned short *)s_q;
    const unsigned char *s_qlo = s_q + 2 * (size_t)K * K;
    const int lane = threadIdx.x;
    {
        const uint4 *src = (const uint4 *)filt;   // (cudaMalloc alignment; 3 K^2 is a multiple of 16 for K >= 4, K = 2 has 12 bytes)
        uint4 *dst = (uint4 *)s_q;
        const int n16 = (3 * K * K) / 16;
        for (int q = lane; q < n16; q += 32) dst[q] = src[q];
        for (int q = n16 * 16 + lane; q < 3 * K * K; q += 32) s_q[q] = filt[q];
    }
    __syncwarp();
    for (int q = lane; q < K; q += 32) {
        W2Replica e;
        e.state = perm_g[q];
        e.diag = w2_image(s_qhi, s_qlo, ((unsigned)q << logK) | (unsigned)e.state);
        e.rowabs = __fmul_ru(1.6e-14f, __fadd_ru(__double2float_ru(filt_rowabs[q]), 0.5f));
        e.pad = 0;
        s_rep[q] = e;
    }
    const unsigned head0 = (unsigned)ctl->head;
    __syncwarp();
    const uint4 *__restrict__ recs = (const uint4 *)rec;

    unsigned h = head0;
    const long long remaining0 = ctl->remaining;
    unsigned rem = remaining0 > 0x7fffffffLL ? 0x7fffffffu : (unsigned)remaining0;
    const unsigned rem0 = rem;
    unsigned rounds = 0, slow = 0;
    // a round may start while h <= h_end: it reads the records of [h, h + 97) (a window, the next window, an advance)
    const unsigned h_end = nslots >= 300u ? nslots - 161u : 0u;
    if (rem >= 130 && nslots >= 300u && h + 99u <= h_end) {
        unsigned r = h & 31u;                           // lane of window position 0
        unsigned w = ((unsigned)lane - h) & 31u;        // this lane's window position
        unsigned sA = h + w;                            // this lane's slot
        // slot contexts: A = current, B = the slot one window later (raw record qB, loaded a round ahead)
        unsigned iA, jA, bmA, iB, jB, bmB;
        float luA, luB;
        auto unpack = [&](const uint4 q, unsigned &i, unsigned &j, unsigned &bm, float &lu) {
            i = q.x & 0xffffu; j = q.x >> 16; bm = q.y; lu = __uint_as_float(q.z);
        };
        unpack(recs[sA], iA, jA, bmA, luA);
        uint4 qB = recs[sA + 32];
        unpack(qB, iB, jB, bmB, luB);
        uint4 ei = *(const uint4 *)&s_rep[iA], ej = *(const uint4 *)&s_rep[jA];   // {state, diag, rowabs} of both replicas
        // what the round resolved last leaves to the next block (predicates are computed where their inputs appear, so
        // that their latency overlaps the loop branch): this lane commits / swaps / leaves the window
        bool p_mine = false, p_swaps = false, p_promoted = false;
        unsigned p_entry = 0, p_advance = 0, p_n = 0;
        float f_ij = 0.f, f_ji = 0.f;
        unsigned z = 0;
        // Commit the round described by the p_* values, slide the window and fetch the states of the next round:
        // permutation stores, at once the next round's loads, then the lane state.
        auto commit = [&]() {
            if (p_swaps) {   // replica i takes state sj: its new diagonal value is the off-diagonal one just read
                *(uint2 *)&s_rep[iA] = make_uint2(ej.x, __float_as_uint(f_ij));
                *(uint2 *)&s_rep[jA] = make_uint2(ei.x, __float_as_uint(f_ji));
            }
            __syncwarp();
            const uint4 eiA = *(const uint4 *)&s_rep[iA], ejA = *(const uint4 *)&s_rep[jA];
            const uint4 eiB = *(const uint4 *)&s_rep[iB], ejB = *(const uint4 *)&s_rep[jB];
            // `z` is the entries' padding word: always zero, but only known once the loads above have returned.  Every
            // piece of book-keeping below is made to depend on it, which keeps the instruction scheduler from issuing it
            // ahead of the stores and loads that head the dependent chain (a warp issues in order).
#if W2_ZDEP
            z = eiA.w;
#endif
            const unsigned adv = p_advance + z;
            if (p_mine) slot_log[sA + z] = p_entry;     // sparse commit log, indexed by slot (zero = no attempt)
            h += adv;
            r = (r + adv) & 31u;
            rem -= p_n + z;
            if (p_promoted) { iA = iB; jA = jB; bmA = bmB; luA = luB; sA += 32; }
            ei = p_promoted ? eiB : eiA;
            ej = p_promoted ? ejB : ejA;
            w = (w - adv) & 31u;
        };
        for (;;) {
            // ---------------- fast rounds.  The loop is rotated: an iteration COMMITS the round resolved by the previous
            // one and then evaluates and resolves the next, so that the block begins with the dependent chain (stores ->
            // state loads -> image loads -> filter -> ballots); no data-dependent branch besides the loop's.
            bool go;
            unsigned cm = 0, adv33 = 0;
            p_mine = p_swaps = p_promoted = false; p_advance = 0; p_n = 0;   // nothing to commit on entry
            do {
                commit();
                rounds++;
                const unsigned si = ei.x, sj = ej.x;
                f_ij = w2_image(s_qhi, s_qlo, (iA << logK) | sj);
                f_ji = w2_image(s_qhi, s_qlo, (jA << logK) | si);
                // the record of the slot one window later: the lanes that stay re-read the one they hold; the lines were
                // brought into L1 two windows ahead and into L2 1024 slots ahead
                const unsigned sZ = sA + z;
                qB = __ldg(recs + (sZ + 32));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(recs + (sZ + 96)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(recs + min(sZ + 1024u, nslots - 1u)));
                // the budget and the end of the pass: the round being resolved now commits at most 32 attempts and
                // advances at most 33 slots, and one more round may follow it before the next test
                go = rem >= 97u && h + 66u <= h_end;
                const float eps0 = (__uint_as_float(ei.z) + __uint_as_float(ej.z)) + 1e-9f;
                bool ge0, acc, undecided;
                w2_filter(__uint_as_float(ei.y), f_ij, __uint_as_float(ej.y), f_ji, eps0, luA, iA == jA, ge0, acc, undecided);
                const unsigned und = undecided ? 1u : 0u;
                const bool changes = acc && iA != jA;
                const unsigned G = __ballot_sync(0xffffffffu, ge0);
                const unsigned A = __ballot_sync(0xffffffffu, changes);
                const unsigned Gw = __funnelshift_r(G, G, r), Aw = __funnelshift_r(A, A, r);   // window order
                // visited chain (see k_mix_walk_pow2): skip[s+1] = NG[s] & ~skip[s]
                const unsigned X = ~Gw;
                const unsigned starts = X & ~(X << 1);
                const unsigned SE = starts & 0x55555555u, SO = starts & 0xAAAAAAAAu;
                const unsigned sumE = X + SE, sumO = X + SO;
                const unsigned skip = (((sumE ^ X) & ~SE) & 0xAAAAAAAAu) | (((sumO ^ X) & ~SO) & 0x55555555u);
                const unsigned V = ~skip;
                const unsigned VA = V & Aw;
                unsigned earlier;   // bit 31-b: window position w-1-b is a visited state-changing swap
                asm("shl.b32 %0, %1, %2;" : "=r"(earlier) : "r"(VA), "r"(32u - w));
                // a round ends before the first visited lane that is stale -- or that the filter could not decide
                const unsigned C = __ballot_sync(0xffffffffu, ((earlier & bmA) | und) != 0u);
                const unsigned Cw = __funnelshift_r(C, C, r) & V;
                const unsigned low = Cw & (0u - Cw);
                const unsigned below = low - 1u;       // low == 0 -> all lanes
                cm = V & below;
                p_mine = (cm >> w) & 1u;
                p_swaps = p_mine && changes;
                p_promoted = (below >> w) & 1u;
                p_entry = si | (sj << LOG_STATE_BITS) | ((acc ? 1u : 0u) << LOG_ACC_BIT) | (1u << 31);
                // bit 32 (even position) of the skip word can only be set by the carry of an odd-start run
                adv33 = (Cw == 0u && sumO < X) ? 1u : 0u;
                p_advance = Cw ? (unsigned)__popc(below) : 32u + adv33;
                p_n = __popc(cm);
                unpack(qB, iB, jB, bmB, luB);   // the next commit's "one window later" context (the load above has landed)
            } while (go && cm != 0u && adv33 == 0u);
            commit();   // the round resolved last
            qB = __ldg(recs + (sA + 32));
            unpack(qB, iB, jB, bmB, luB);
            // ---------------- rare events
            if (adv33) {
                // the slot after the window is a uniform's slot: the lane that is now at window position 31 holds slot
                // h - 1 (skipped) and moves on by another window
                if (w == 31u) {
                    sA += 32;
                    unpack(recs[sA], iA, jA, bmA, luA);
                    qB = recs[sA + 32];
                    unpack(qB, iB, jB, bmB, luB);
                    ei = *(const uint4 *)&s_rep[iA]; ej = *(const uint4 *)&s_rep[jA];
                }
                __syncwarp();
            }
            if (rem < 130u || h + 99u > h_end) break;
            if (cm == 0u) {
                // the first lane of the window is undecided: one exact attempt, exactly what the reference does
                rounds++;
                bool ge0 = false, acc = false;
                const unsigned si = ei.x, sj = ej.x;
                f_ij = w2_image(s_qhi, s_qlo, (iA << logK) | sj);
                f_ji = w2_image(s_qhi, s_qlo, (jA << logK) | si);
                if (w == 0u) {
                    const unsigned rowi = iA << logK, rowj = jA << logK;
                    const double logp = swap_logp(u[rowi | sj], u[rowj | si], u[rowi | si], u[rowj | sj]);
                    ge0 = logp >= 0.0;
                    acc = ge0;
                    if (!ge0) {
                        const unsigned s1 = sA + 1;
                        const double dd = logp - slot_logU(words, s1);
                        if (dd > 1e-9) acc = true;
                        else if (dd < -1e-9) acc = false;
                        else acc = mt_double(words[2 * (size_t)s1], words[2 * (size_t)s1 + 1]) < exp(logp);
                    }
                    slow++;
                }
                __syncwarp();
                const bool first_ge0 = __ballot_sync(0xffffffffu, ge0) != 0u;
                p_advance = first_ge0 ? 1u : 2u;   // log_p < 0: the next slot is this attempt's uniform
                p_n = 1u;
                p_mine = (w == 0u);
                p_swaps = p_mine && acc && iA != jA;
                p_promoted = w < p_advance;
                p_entry = si | (sj << LOG_STATE_BITS) | ((acc ? 1u : 0u) << LOG_ACC_BIT) | (1u << 31);
                commit();
                qB = __ldg(recs + (sA + 32));
                unpack(qB, iB, jB, bmB, luB);
                if (rem < 130u || h + 99u > h_end) break;
            }
        }
        __syncwarp();
        for (int q = lane; q < K; q += 32) perm_g[q] = s_rep[q].state;
    }
    slow = __reduce_add_sync(0xffffffffu, slow);
    if (lane == 0) {
        const long long remaining = remaining0 - (long long)(rem0 - rem);
        ctl->head = h;
        ctl->remaining = remaining;
        ctl->status = remaining > 0 ? 1 : 0;
        ctl->rounds += (int)rounds;
        ctl->slow_exp += slow;
    }
}

// Count matrices from the sparse commit log of k_mix_walk2 (one word per slot, bit 31 = an attempt started there).
__global__ void k_mix_count_slots(const uint32_t *__restrict__ slot_log, long long s0, long long s1, int M,
                                  unsigned long long *__restrict__ nacc, unsigned long long *__restrict__ nprop) {
    long long t = s0 + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < s1; t += stride) {
        const uint32_t e = slot_log[t];
        if (!(e >> 31)) continue;
        const uint32_t si = e & ((1u << LOG_STATE_BITS) - 1u), sj = (e >> LOG_STATE_BITS) & ((1u << LOG_STATE_BITS) - 1u);
        atomicAdd(&nprop[(size_t)si * M + sj], 1ull);
        atomicAdd(&nprop[(size_t)sj * M + si], 1ull);
        if ((e >> LOG_ACC_BIT) & 1u) {
            atomicAdd(&nacc[(size_t)si * M + sj], 1ull);
            atomicAdd(&nacc[(size_t)sj * M + si], 1ull);
        }
    }
}
