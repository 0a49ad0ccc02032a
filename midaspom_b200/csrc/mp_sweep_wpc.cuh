// mp_sweep_wpc.cuh -- the FP32 Gibbs scan of y with a WINDOW of candidates in flight, one warp (or a few) per candidate.
//
// Same update, same visiting order, same decisions as k_sweep_y_cull (mp_sweep_cull.cuh); what changes is who does what.
// There every trip spreads ONE pass over all targets over all warps of the task and pays a task-wide reduction, so a trip
// costs ~860 instructions per warp of which ~180 evaluate weights.  Here:
//
//  * one CTA (16 warps) owns the task; the targets sit in shared memory GROUP-MAJOR (group = 32 consecutive slots of the
//    layout order = 32 spatial neighbours, lane l <-> slot 32 g + l), with per-group bounds {centre, radius, min S} and
//    class masks next to them, and per-supergroup (32 groups) bounds above those;
//  * every round evaluates a window of W consecutive candidates against the CURRENT state, candidate c by its own WPC
//    warps: lane-parallel bound tests find the groups that matter, the warp evaluates them 32 targets at a time, and the
//    only cross-warp traffic of the round is one (sum, remainder bound) pair per warp;
//  * EARLY DECISION with a rigorous remainder bound.  A group is evaluated when the candidate's largest possible weight
//    on it exceeds 2^-9 of the group's smallest S (NEAR).  For the groups between that and the FP32 resolution 2^-26
//    (MID; below it the factor ratio is exactly 1, as in k_sweep_y_cull) the warp only adds an upper bound of their
//    |log-ratio| to R:   n_A w/(S_min - w) + n_B cK w/(slack_min - cK w)   (classes A: z'=1, factor cK S; B: z'=0, factor
//    1 - cK S; slack = 1 - cK S).  If the threshold lies outside [L - R, L + R] the flip is decided exactly as the full sum
//    would decide it; otherwise (about 1 % of the candidates) all warps evaluate the MID groups and the full sum decides;
//  * the candidates of the window are decided in order; the first accepted (or undecided) one ends the round: an accepted
//    flip is committed by all warps (groups interleaved over the warps, same 2^-36 commit rule and two-float S as before)
//    and the candidates behind it are evaluated again next round.  2 of 3 flips are rejections, so a round retires
//    about 3 candidates.
//
// Decisions equal those of the one-at-a-time scan up to FP32 ties (tests: the FP64 engine flip by flip).
#pragma once
#include "mp_sweep_cull.cuh"

namespace mp {

constexpr int WPC_NT = 512, WPC_NW = WPC_NT / 32;
constexpr float WPC_FEEL = 1.4886e-8f;            // 2^-26 with the margin of k_sweep_y_cull: below it a group contributes exactly 0
constexpr float WPC_COMMIT = 1.4551915e-11f;      // 2^-36: commit rule of k_sweep_y_cull
constexpr float WPC_FULLC = 9.5367432e-7f;        // 2^-20: commits below it touch only S_lo (no renormalisation, no new group minimum)

// shared memory of a task with at most nl_max targets (host and device agree through this one function)
__host__ __device__ inline size_t wpc_smem_bytes(int nl_max)
{
    const size_t G = (size_t)((((nl_max + 31) / 32) + 3) & ~3);
    return G * 32 * 16 + G * (16 + 4 * 4 + 4) + 16 * 16 + 128 * 16 + 2 * 16 * 8 + 16 * 4 + 16 * 16 * 4 + 64;
}

template <int GEOM, int W, int WPC>
__global__ void __launch_bounds__(WPC_NT, 1)
k_sweep_y_wpc(Landscape<float> ls, const int *__restrict__ perm, const mp_params *__restrict__ par,
              const uint8_t *__restrict__ era, const uint8_t *__restrict__ z, uint8_t *__restrict__ y, double *__restrict__ S,
              const CandRec *__restrict__ rec, const int *__restrict__ count, int T, int nl_max, const int *__restrict__ order,
              unsigned long long *__restrict__ stats, const BlockTask *__restrict__ btasks, const int *__restrict__ tlist,
              const int *__restrict__ blockmode, float near_thr /* groups evaluated in the first pass: largest weight >= near_thr x the group's min S ... */,
              float bcap /* ... or a remainder bound above bcap (log2 units) */)
{
    static_assert(GEOM != MP_GEOM_DENSE, "needs positions");
    static_assert(W * WPC == WPC_NW, "W candidates x WPC warps each = the CTA's 16 warps");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = ls.n, ntrans = T - 1, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const BlockTask bt = btasks ? btasks[blockIdx.x] : BlockTask{ order[blockIdx.x], 0, n, 0, n };
    const int task = bt.task;
    const int c = task / ntrans, t = task - c * ntrans;
    if (blockmode && (blockmode[task] != 0) != (btasks != nullptr)) return;
    const int nl = bt.nl;
    const int *tl = btasks ? tlist + bt.tl_off : perm;
    const int G = (nl + 31) / 32, Gp = (((nl_max + 31) / 32) + 3) & ~3, SG = (G + 31) / 32;
    float4 *sT = reinterpret_cast<float4 *>(smem_raw);               // {S_hi, S_lo, x, y} (COORDS) / {S_hi, S_lo, patch, -} (LINEAR), slot 32 g + lane
    float4 *gbox = sT + (size_t)Gp * 32;                             // per group: COORDS {cx, cy, radius, min S_hi}; LINEAR {first, last patch, 0, min S_hi}
    uint32_t *gA = reinterpret_cast<uint32_t *>(gbox + Gp);          // class A lanes (y=0, z'=1)
    uint32_t *gB = gA + Gp, *gV = gB + Gp, *gY = gV + Gp;            // class B lanes (y=0, z'=0); valid lanes; y bits
    float *gsl = reinterpret_cast<float *>(gY + Gp);                 // min over the class-B lanes of 1 - cK S (lower bound)
    float4 *sbox = reinterpret_cast<float4 *>(gsl + Gp);             // per supergroup of 32 groups: the same bounds
    float4 *ring = sbox + 16;                                        // 2 chunks of 32 candidate records (two float4 each)
    float2 *res = reinterpret_cast<float2 *>(ring + 128);            // [round parity][warp] {partial sum, partial remainder bound}
    float *red = reinterpret_cast<float *>(res + 2 * WPC_NW);        // [warp] partial sums of the second pass
    uint32_t *nmask = reinterpret_cast<uint32_t *>(red + WPC_NW);    // [candidate of the window][supergroup] NEAR groups found by the candidate's warps

    const Trans<float> tr = make_trans<float>(par[c], era ? era[t] : 0);
    const float cK = tr.c * tr.Kt;
    const float nal2e = alpha_pre<float>(par[c].alpha);
    const double src_scale = tr.src ? (double)tr.Ks / (double)tr.Kt : 0.0;
    auto src_offset = [&](int q) -> double {
        if (!tr.src) return 0.0;
        const double u = ls.src_unit ? ls.src_unit[q] : (double)(q + 1);
        return src_scale * exp(-(par[c].alpha * u) * par[c].dsrc);
    };
    uint8_t *yt = y + ((size_t)c * ntrans + t) * n;
    const uint8_t *zn = z + ((size_t)c * T + t + 1) * n;
    double *St = S + ((size_t)c * ntrans + t) * n;
    const float BIG = 3.0e38f;

    // ---- load the targets group by group and form the group bounds
    for (int g = wid; g < G; g += WPC_NW) {
        const int s = g * 32 + lane;
        const int q = s < nl ? tl[s] : -1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t yq = 0, zq = 0;
        if (q >= 0) {
            if (GEOM == MP_GEOM_COORDS) { v.z = ls.px[q]; v.w = ls.py[q]; }
            else v.z = __int_as_float(q);
            const double sv = St[q] + src_offset(q);
            v.x = (float)sv; v.y = (float)(sv - (double)v.x);
            yq = yt[q] != 0; zq = zn[q] != 0;
        }
        sT[s] = v;
        const bool va = q >= 0, ca = va && !yq && zq, cb = va && !yq && !zq;
        const uint32_t Am = __ballot_sync(0xffffffffu, ca), Bm = __ballot_sync(0xffffffffu, cb);
        const uint32_t Vm = __ballot_sync(0xffffffffu, va), Ym = __ballot_sync(0xffffffffu, va && yq);
        float a0, a1, b0, b1, ml = va ? v.x : BIG, sl = cb ? fmaxf(1.f - cK * (v.x + v.y), 0.f) : BIG;
        if (GEOM == MP_GEOM_COORDS) { a0 = va ? v.z : BIG; a1 = va ? v.z : -BIG; b0 = va ? v.w : BIG; b1 = va ? v.w : -BIG; }
        else { a0 = va ? (float)q : BIG; a1 = va ? (float)q : -BIG; b0 = 0.f; b1 = 0.f; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 = fminf(a0, __shfl_xor_sync(0xffffffffu, a0, o)); a1 = fmaxf(a1, __shfl_xor_sync(0xffffffffu, a1, o));
            b0 = fminf(b0, __shfl_xor_sync(0xffffffffu, b0, o)); b1 = fmaxf(b1, __shfl_xor_sync(0xffffffffu, b1, o));
            ml = fminf(ml, __shfl_xor_sync(0xffffffffu, ml, o)); sl = fminf(sl, __shfl_xor_sync(0xffffffffu, sl, o));
        }
        if (lane == 0) {
            if (GEOM == MP_GEOM_COORDS) { const float dx = a1 - a0, dy = b1 - b0;
                                          gbox[g] = make_float4(0.5f * (a0 + a1), 0.5f * (b0 + b1), 0.5000005f * sqrtf(dx * dx + dy * dy) + 1e-3f, ml); }
            else gbox[g] = make_float4(a0, a1, 0.f, ml);
            gA[g] = Am; gB[g] = Bm; gV[g] = Vm; gY[g] = Ym; gsl[g] = sl;
        }
    }
    __syncthreads();
    // supergroup s covers groups 32 s .. 32 s + 31: a circle around their circles (COORDS) / their patch range (LINEAR), min of their min S
    if (wid < SG) {
        const int g = wid * 32 + lane;
        const bool va = g < G;
        const float4 b = va ? gbox[g] : make_float4(0.f, 0.f, 0.f, BIG);
        float a0, a1, b0, b1, ml = b.w;
        if (GEOM == MP_GEOM_COORDS) { a0 = va ? b.x - b.z : BIG; a1 = va ? b.x + b.z : -BIG; b0 = va ? b.y - b.z : BIG; b1 = va ? b.y + b.z : -BIG; }
        else { a0 = va ? b.x : BIG; a1 = va ? b.y : -BIG; b0 = 0.f; b1 = 0.f; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 = fminf(a0, __shfl_xor_sync(0xffffffffu, a0, o)); a1 = fmaxf(a1, __shfl_xor_sync(0xffffffffu, a1, o));
            b0 = fminf(b0, __shfl_xor_sync(0xffffffffu, b0, o)); b1 = fmaxf(b1, __shfl_xor_sync(0xffffffffu, b1, o));
            ml = fminf(ml, __shfl_xor_sync(0xffffffffu, ml, o));
        }
        if (lane == 0) {
            if (GEOM == MP_GEOM_COORDS) { const float dx = a1 - a0, dy = b1 - b0;
                                          sbox[wid] = make_float4(0.5f * (a0 + a1), 0.5f * (b0 + b1), 0.5000005f * sqrtf(dx * dx + dy * dy) + 1e-3f, ml); }
            else sbox[wid] = make_float4(a0, a1, 0.f, ml);
        }
    }

    // ---- candidates (visiting order); a block owns the run whose positions lie in [slot_lo, slot_hi)
    const CandRec *recs = rec + (size_t)task * n;
    int ncand = count[2 * task];
    int nocc = count[2 * task + 1];
    if (btasks) {
        auto first_at_least = [&](int slot) {
            int lo = 0, hi = ncand;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)recs[mid].pad[1] < slot) lo = mid + 1; else hi = mid; }
            return lo;
        };
        const int i0 = first_at_least(bt.slot_lo), i1 = first_at_least(bt.slot_hi);
        recs += i0; ncand = i1 - i0;
        nocc = 1 << 30;      // other blocks change the year's count concurrently; an empty year cannot arise (k_block_valid)
    }
    const int slot_lo = bt.slot_lo;
    const bool blk = btasks != nullptr;
    constexpr int RC = 32;
    const float4 *recs4 = reinterpret_cast<const float4 *>(recs);
    auto load_chunk = [&](int cidx) -> float4 {
        const int idx = cidx * RC * 2 + tid;
        return (tid < 2 * RC && idx < 2 * ncand) ? recs4[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 pre;
    {
        const float4 c0 = load_chunk(0), c1 = load_chunk(1);
        pre = load_chunk(2);
        if (tid < 2 * RC) { ring[tid] = c0; ring[2 * RC + tid] = c1; }
    }
    __syncthreads();
    int chunk = 0;

    struct Cand { int k, kg, kl; uint32_t cur; float kx, ky, lawk, thr, sgn; };
    auto read_cand = [&](int i) -> Cand {
        const float4 a = ring[(i & (2 * RC - 1)) * 2], b = ring[(i & (2 * RC - 1)) * 2 + 1];
        Cand cd;
        cd.k = (int)(__float_as_uint(a.x) & 0x7fffffffu); cd.cur = __float_as_uint(a.x) >> 31;
        cd.kx = a.y; cd.ky = a.z; cd.lawk = a.w; cd.thr = b.x; cd.sgn = cd.cur ? -1.f : 1.f;
        const int ks = blk ? (int)__float_as_uint(b.z) - slot_lo : (int)__float_as_uint(b.y);   // slot of the candidate among the targets
        cd.kg = ks >> 5; cd.kl = ks & 31;
        return cd;
    };
    // largest possible weight of the candidate on anything inside a bound {centre, radius} / {first, last patch}
    auto bound = [&](const float4 &b, const Cand &cd) -> float {
        float dlow;
        if (GEOM == MP_GEOM_COORDS) { const float dx = b.x - cd.kx, dy = b.y - cd.ky; dlow = fmaxf(0.f, Num<float>::sqrtv(fmaf(dx, dx, dy * dy)) - b.z); }
        else dlow = ls.spacing * fmaxf(0.f, fmaxf(b.x - (float)cd.k, (float)cd.k - b.y));
        return 1.001f * weight_of(nal2e, cd.lawk, 0.99999f * dlow);
    };
    auto weight = [&](const float4 &tq, const Cand &cd) -> float {
        float d;
        if (GEOM == MP_GEOM_LINEAR) { const int q = __float_as_int(tq.z); d = (float)(q > cd.k ? q - cd.k : cd.k - q) * ls.spacing; }
        else { const float dx = tq.z - cd.kx, dy = tq.w - cd.ky; d = Num<float>::sqrtv(fmaf(dx, dx, dy * dy)); }
        return weight_of(nal2e, cd.lawk, d);
    };
    // class of group g for the candidate: 0 out of reach (contributes exactly 0), 1 MID (bounded by `bnd`), 2 NEAR (evaluated)
    auto classify = [&](int g, const Cand &cd, float &bnd) -> int {
        bnd = 0.f;
        if (g >= G) return 0;
        const float4 b = gbox[g];
        const float wub = bound(b, cd), ml = b.w;
        if (wub < WPC_FEEL * ml) return 0;
        if (!(wub < near_thr * ml)) return 2;
        const float nA = (float)__popc(gA[g]), nB = (float)__popc(gB[g]);
        float v = nA * __fdividef(wub, ml - wub);
        if (nB > 0.f) { const float den = gsl[g] - cK * wub; v = den > 0.f ? v + nB * __fdividef(cK * wub, den) : BIG; }
        v *= 1.4426950408889634f * 1.02f;                             // log2 units, margin for the approximate divisions
        if (!(v < bcap)) return 2;
        bnd = v;
        return 1;
    };
    // sum over the groups {base + stride * bit : bit in m} of lg2(new factor / current factor), this lane's target of each
    auto eval_groups = [&](uint32_t m, int base, int stride, const Cand &cd) -> float {
        float acc = 0.f, pn = 1.f, pd = 1.f;
        int cnt = 0;
        auto batch = [&](auto BB) {
            constexpr int B = decltype(BB)::value;
            int gg[B]; float4 tq[B]; uint32_t am[B], bm[B];
#pragma unroll
            for (int u = 0; u < B; u++) {
                gg[u] = base + stride * (__ffs((int)m) - 1); m &= m - 1u;
                tq[u] = sT[gg[u] * 32 + lane]; am[u] = gA[gg[u]]; bm[u] = gB[gg[u]];
            }
#pragma unroll
            for (int u = 0; u < B; u++) {
                const bool a = (am[u] >> lane) & 1u, b = (bm[u] >> lane) & 1u;
                const float mul = a ? cK : (b ? -cK : 0.f), add = a ? 0.f : 1.f;
                const float fd = __saturatef(fmaf(mul, tq[u].x + tq[u].y, add));
                const float w = weight(tq[u], cd);
                const float fn = __saturatef(fmaf(mul, fmaf(cd.sgn, w, tq[u].x) + tq[u].y, add));
                pd *= fd;
                pn *= (gg[u] == cd.kg && lane == cd.kl) ? fd : fn;       // the candidate's own cell is handled by own_term
            }
            cnt += B;
        };
        while (m) {                                                     // warp-uniform
            const int left = __popc(m);
            if (left >= 4) batch(std::integral_constant<int, 4>());
            else if (left == 3) batch(std::integral_constant<int, 3>());
            else if (left == 2) batch(std::integral_constant<int, 2>());
            else batch(std::integral_constant<int, 1>());
            if (cnt >= 4 || m == 0u) { acc += Num<float>::lg2(pn) - Num<float>::lg2(pd); pn = 1.f; pd = 1.f; cnt = 0; }
        }
        return acc;
    };
    auto own_term = [&](const Cand &cd) -> float {
        const float4 tq = sT[cd.kg * 32 + cd.kl];
        const float lg2e = 1.4426950408889634f;
        const float l0 = lg2e * tr.logE + Num<float>::lg2(__saturatef(cK * (tq.x + tq.y)));
        const float l1 = lg2e * tr.log1mE;
        return cd.cur ? ldiff<float>(l0, l1) : ldiff<float>(l1, l0);
    };
    // block sum of one float per warp partial (fixed order), one barrier
    auto block_sum16 = [&](float v) -> float {
        v = warp_sum_f(v);
        if (lane == 0) red[wid] = v;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < WPC_NW; w++) tot += red[w];
        __syncthreads();
        return tot;
    };

    uint32_t st_trips = 0, st_exec = 0, st_ret = 0, st_commit = 0;
#ifdef MP_WPC_PROFILE
    long long pf_t[5] = { 0, 0, 0, 0, 0 }, pf_mark = clock64();         // evaluation, decision, second pass, commit, other
    unsigned long long pf_n[5] = { 0, 0, 0, 0, 0 };                     // rounds, second passes, careful passes, commits, retired candidates
#define PF_LAP(slot) do { const long long now_ = clock64(); pf_t[slot] += now_ - pf_mark; pf_mark = now_; } while (0)
#else
#define PF_LAP(slot) do { } while (0)
#endif
    const int cw = wid / WPC, part = wid % WPC;                        // this warp's candidate of the window and its share of it
    uint32_t round = 0;
    for (int i = 0; i < ncand; round++) {
        if ((i / RC) != chunk) {                                       // entered the next chunk: recycle the half we left
            chunk = i / RC;
            __syncthreads();
            if (tid < 2 * RC) ring[((chunk + 1) & 1) * 2 * RC + tid] = pre;
            __syncthreads();
            pre = load_chunk(chunk + 2);
        }
        const int nc = min(W, ncand - i);
        PF_LAP(4);
        // ---- evaluation: candidate i + cw by its WPC warps.  The supergroups are classified round-robin by the parts, the
        // NEAR masks meet in shared memory, then part p evaluates the NEAR groups g with g mod WPC == p (balanced: NEAR
        // groups come in runs of consecutive layout slots)
        uint32_t my_exec = 0;
        {
            float acc = 0.f, Rb = 0.f;
            const Cand cd = read_cand(i + cw);
            const bool zero = (nocc + (cd.cur ? -1 : 1)) == 0;
            if (cw < nc && !zero) {
                bool sfar = true;
                if (lane < SG) { const float4 sb = sbox[lane]; sfar = bound(sb, cd) < WPC_FEEL * sb.w; }
                const uint32_t sm = __ballot_sync(0xffffffffu, !sfar);
                uint32_t word = 0;                                      // lane sg: NEAR groups of supergroup sg
                {
                    uint32_t rest = sm;
                    for (int idx = 0; rest; idx++) {                     // warp-uniform
                        const int sg = __ffs((int)rest) - 1; rest &= rest - 1u;
                        if (WPC > 1 && (idx % WPC) != part) continue;
                        float bnd;
                        const int cls = classify(sg * 32 + lane, cd, bnd);
                        Rb += bnd;
                        const uint32_t nm = __ballot_sync(0xffffffffu, cls == 2);
                        if (WPC > 1) { if (lane == 0) nmask[cw * 16 + sg] = nm; }
                        else if (lane == sg) word = nm;
                    }
                }
                if (WPC > 1) {
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + cw), "r"(WPC * 32) : "memory");
                    word = (lane < 16 && ((sm >> lane) & 1u)) ? nmask[cw * 16 + lane] : 0u;
                }
                constexpr uint32_t PARTMASK = WPC == 1 ? 0xffffffffu : WPC == 2 ? 0x55555555u : 0x11111111u;
                uint32_t rest = sm;
                while (rest) {
                    const int sg = __ffs((int)rest) - 1; rest &= rest - 1u;
                    const uint32_t nm = __shfl_sync(0xffffffffu, word, sg) & (PARTMASK << part);
                    my_exec += __popc(nm);
                    acc += eval_groups(nm, sg * 32, 1, cd);
                }
                if (part == 0 && lane == 0) acc += own_term(cd);
            }
            acc = warp_sum_f(acc); Rb = warp_sum_f(Rb);
            if (lane == 0) res[(round & 1u) * WPC_NW + wid] = make_float2(acc, Rb);
        }
        __syncthreads();
        PF_LAP(0);
        // ---- decisions, in order; the first accepted or undecided candidate ends the round
        int adv = nc, verdict = 0;                                      // 1: accept (commit), 2: undecided by the bound (second pass)
        float tot_near = 0.f;
        {   // lane cc judges candidate i + cc; the first one that is not a clean rejection ends the round
            int v = 0;
            float tot = 0.f;
            if (lane < nc) {
                float R = 0.f;
#pragma unroll
                for (int p = 0; p < WPC; p++) { const float2 r = res[(round & 1u) * WPC_NW + lane * WPC + p]; tot += r.x; R += r.y; }
                const float4 ra = ring[((i + lane) & (2 * RC - 1)) * 2], rb = ring[((i + lane) & (2 * RC - 1)) * 2 + 1];
                const bool zero = (nocc + ((__float_as_uint(ra.x) >> 31) ? -1 : 1)) == 0;
                const float L = 0.6931471805599453f * tot, mR = 0.6931471805599453f * R, thr = rb.x;
                if (zero || isnan(tot) || !(mR < BIG)) v = 2;
                else if (thr < L - mR) v = 1;
                else if (!(thr >= L + mR)) v = 2;
            }
            const uint32_t stop = __ballot_sync(0xffffffffu, v != 0);
            if (stop) {
                const int first = __ffs((int)stop) - 1;
                adv = first + 1;
                verdict = __shfl_sync(0xffffffffu, v, first);
                tot_near = __shfl_sync(0xffffffffu, tot, first);
            }
        }
        st_trips++; st_exec += my_exec; if (cw < adv) st_ret += my_exec;
        PF_LAP(1);
#ifdef MP_WPC_PROFILE
        pf_n[0]++; pf_n[4] += adv;
#endif
        if (verdict) {
            const Cand cd = read_cand(i + adv - 1);
            const bool zero = (nocc + (cd.cur ? -1 : 1)) == 0;
            if (verdict == 2) {
                float tot;
#ifdef MP_WPC_PROFILE
                pf_n[1]++; if (zero || isnan(tot_near)) pf_n[2]++;
#endif
                if (zero || isnan(tot_near)) {
                    // cell by cell with the (-inf) - (-inf) := 0 convention (impossible current state, removal of the last occupied patch)
                    float acc2 = 0.f;
                    for (int g = wid; g < G; g += WPC_NW) {
                        const float4 tq = sT[g * 32 + lane];
                        const bool own = g == cd.kg && lane == cd.kl;
                        const bool a = ((gA[g] >> lane) & 1u) && !own, b = (gB[g] >> lane) & 1u;
                        if (!(a || b)) continue;
                        const float w = weight(tq, cd);
                        const int s = g * 32 + lane;
                        const float sa = zero ? (float)src_offset(tl[s]) : fmaf(cd.sgn, w, tq.x) + tq.y;
                        acc2 += ldiff<float>(Num<float>::lg2(col_factor(cK, sa, a, b)), Num<float>::lg2(col_factor(cK, tq.x + tq.y, a, b)));
                    }
                    if (tid == 0) acc2 += own_term(cd);
                    tot = block_sum16(acc2);
                } else {
                    // second pass: the MID groups, interleaved over the warps (lane l tests group wid + 16 l)
                    float bnd;
                    const int cls = classify(wid + WPC_NW * lane, cd, bnd);
                    const uint32_t mm = __ballot_sync(0xffffffffu, cls == 1);
                    st_exec += __popc(mm); st_ret += __popc(mm);
                    tot = tot_near + block_sum16(eval_groups(mm, wid, WPC_NW, cd));
                }
                verdict = cd.thr < 0.6931471805599453f * tot ? 1 : 0;
                PF_LAP(2);
            }
            if (verdict == 1) {
                // ---- commit: rank-1 update of S on every group within the commit reach, groups interleaved over the warps
                if (!zero) {
                    const int gt = wid + WPC_NW * lane;
                    bool reach = false, full = false;
                    float wubg = 0.f;
                    if (gt < G) { const float4 b = gbox[gt]; wubg = bound(b, cd); reach = !(wubg < WPC_COMMIT * b.w); full = !(wubg < WPC_FULLC * b.w); }
                    uint32_t m = __ballot_sync(0xffffffffu, reach && full), mlo = __ballot_sync(0xffffffffu, reach && !full);
                    st_commit += __popc(m) + __popc(mlo);
                    // an addition lowers the slack 1 - cK S of the far groups by at most cK x the largest weight: keep the bound valid
                    if (reach && !full && !cd.cur) gsl[gt] = fmaxf(gsl[gt] - cK * wubg, 0.f);
                    auto batch = [&](auto BB) {
                        constexpr int B = decltype(BB)::value;
                        int gg[B]; float4 tq[B];
#pragma unroll
                        for (int u = 0; u < B; u++) { gg[u] = wid + WPC_NW * (__ffs((int)m) - 1); m &= m - 1u; tq[u] = sT[gg[u] * 32 + lane]; }
#pragma unroll
                        for (int u = 0; u < B; u++) {
                            const float w = weight(tq[u], cd);
                            const float a = (gg[u] == cd.kg && lane == cd.kl) ? 0.f : cd.sgn * w;
                            const float s = tq[u].x + a, bb = s - tq[u].x;
                            const float e = (tq[u].x - (s - bb)) + (a - bb);
                            const float lo2 = tq[u].y + e;
                            float hi = s + lo2, lo = lo2 - (hi - s);
                            if (hi < 0.f) { hi = 0.f; lo = 0.f; }
                            *reinterpret_cast<float2 *>(&sT[gg[u] * 32 + lane]) = make_float2(hi, lo);
                            if (cd.cur) {
                                // a removal lowers S: exact new minimum of the group (S_hi >= 0: the bit pattern orders like the value)
                                const uint32_t mn = __reduce_min_sync(0xffffffffu, ((gV[gg[u]] >> lane) & 1u) ? (__float_as_uint(hi) & 0x7fffffffu) : 0x7f7fffffu);
                                if (lane == 0) { gbox[gg[u]].w = __uint_as_float(mn); atomicMin(reinterpret_cast<unsigned int *>(&sbox[gg[u] >> 5].w), mn); }
                            } else {
                                // an addition lowers the slack 1 - cK S of the class-B cells: exact new minimum
                                const float sl = fmaxf(1.f - cK * (hi + lo), 0.f);
                                const uint32_t mn = __reduce_min_sync(0xffffffffu, ((gB[gg[u]] >> lane) & 1u) ? __float_as_uint(sl) : 0x7f7fffffu);
                                if (lane == 0) gsl[gg[u]] = __uint_as_float(mn);
                            }
                        }
                    };
                    while (m) {
                        const int left = __popc(m);
                        if (left >= 4) batch(std::integral_constant<int, 4>());
                        else if (left == 3) batch(std::integral_constant<int, 3>());
                        else if (left == 2) batch(std::integral_constant<int, 2>());
                        else batch(std::integral_constant<int, 1>());
                    }
                    // far groups: the weight is below 2^-20 of every S of the group -- it goes into S_lo alone (S_hi and with it the
                    // group minimum stay; every evaluation reads S_hi + S_lo; the next full commit of the group renormalises the pair)
                    auto batch_lo = [&](auto BB) {
                        constexpr int B = decltype(BB)::value;
                        int gg[B]; float4 tq[B];
#pragma unroll
                        for (int u = 0; u < B; u++) { gg[u] = wid + WPC_NW * (__ffs((int)mlo) - 1); mlo &= mlo - 1u; tq[u] = sT[gg[u] * 32 + lane]; }
#pragma unroll
                        for (int u = 0; u < B; u++) sT[gg[u] * 32 + lane].y = fmaf(cd.sgn, weight(tq[u], cd), tq[u].y);
                    };
                    while (mlo) {
                        const int left = __popc(mlo);
                        if (left >= 4) batch_lo(std::integral_constant<int, 4>());
                        else if (left == 3) batch_lo(std::integral_constant<int, 3>());
                        else if (left == 2) batch_lo(std::integral_constant<int, 2>());
                        else batch_lo(std::integral_constant<int, 1>());
                    }
                } else {
                    // the last occupied patch leaves: S is exactly the external source term everywhere
                    for (int g = wid; g < G; g += WPC_NW) {
                        const int s = g * 32 + lane;
                        const double so = s < nl ? src_offset(tl[s]) : 0.0;
                        const float hi = (float)so, lo = (float)(so - (double)hi);
                        *reinterpret_cast<float2 *>(&sT[s]) = make_float2(hi, lo);
                        if (lane == 0) { gbox[g].w = 0.f; gsl[g] = 0.f; }
                    }
                    if (tid < 16) sbox[tid].w = 0.f;
                }
                if (tid == 0) { gY[cd.kg] ^= 1u << cd.kl; gA[cd.kg] ^= 1u << cd.kl; }   // y flips; y=1 cells carry no colonisation factor
                nocc += cd.cur ? -1 : 1;
                __syncthreads();
                PF_LAP(3);
#ifdef MP_WPC_PROFILE
                pf_n[3]++;
#endif
            }
        }
        i += adv;
    }
    __syncthreads();
    for (int g = wid; g < G; g += WPC_NW) {
        const int s = g * 32 + lane;
        if (s < nl) {
            const int q = tl[s];
            const float4 tq = sT[s];
            St[q] = fmax(((double)tq.x + (double)tq.y) - src_offset(q), 0.0);
            yt[q] = (uint8_t)((gY[g] >> lane) & 1u);
        }
    }
    if (stats && lane == 0) {
        atomicAdd(&stats[MP_CNT_SCAN_TRIPS], (unsigned long long)st_trips);
        atomicAdd(&stats[MP_CNT_SCAN_EXEC], (unsigned long long)st_exec);
        atomicAdd(&stats[MP_CNT_SCAN_RETIRED], (unsigned long long)st_ret);
        atomicAdd(&stats[MP_CNT_SCAN_COMMIT], (unsigned long long)st_commit);
        if (btasks && tid == 0) atomicAdd(&stats[MP_CNT_SCAN_BLOCKS], 1ull);
    }
#ifdef MP_WPC_PROFILE
    if (stats && tid == 0) {
        for (int q = 0; q < 5; q++) { atomicAdd(&stats[MP_CNT_N + q], (unsigned long long)pf_t[q]); atomicAdd(&stats[MP_CNT_N + 5 + q], pf_n[q]); }
    }
#endif
}

}  // namespace mp

#ifdef MP_WPC_GEOM
#include "mp_host.h"
namespace mp {
template <int W, int WPC> static int launch_wpc(mp_engine *h, int nl_max, int nclusters, const BlockTask *btasks)
{
    const size_t smem = wpc_smem_bytes(nl_max);
    REQUIRE(smem <= 227 * 1024, MP_ERR_UNSUPPORTED, "too many targets for the one-CTA y scan");
    auto kern = k_sweep_y_wpc<MP_WPC_GEOM, W, WPC>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (nclusters <= 0) return MP_OK;
    kern<<<nclusters, WPC_NT, smem, h->stream>>>(view<float>(h), (const int *)h->d_perm, (const mp_params *)h->d_par,
                                                 (const uint8_t *)(h->have_era ? h->d_era : nullptr), (const uint8_t *)h->d_z, h->d_y, h->d_S[0],
                                                 (const CandRec *)h->d_cand, (const int *)h->d_cand_count, h->cfg.n_years, nl_max,
                                                 (const int *)h->d_task_order, h->d_work, btasks, (const int *)h->d_tlist,
                                                 (const int *)(h->blk_active ? h->d_blockmode : nullptr), h->wpc_near, h->wpc_bcap);
    CK(cudaGetLastError());
    return MP_OK;
}
// window: candidates in flight per round (16 / 8 / 4, with 1 / 2 / 4 warps each)
static int launch_wpc_any(mp_engine *h, int window, int nl_max, int nclusters, const void *btasks_v)
{
    const BlockTask *btasks = (const BlockTask *)btasks_v;
    if (window == 16) return launch_wpc<16, 1>(h, nl_max, nclusters, btasks);
    if (window == 4) return launch_wpc<4, 4>(h, nl_max, nclusters, btasks);
    return launch_wpc<8, 2>(h, nl_max, nclusters, btasks);
}
}  // namespace mp
#endif
