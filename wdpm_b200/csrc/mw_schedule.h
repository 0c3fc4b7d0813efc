// Schedule of the fused ("marching window") iteration kernel.
//
// One CTA owns a strip of TWV grid columns and a chunk of rows. It stages a
// window of W columns (left halo HL + TWV owned + right halo) in shared memory and
// marches down the rows three at a time ("row triples"). With padded row r and
// the reference's colour offsets oi,oj in 1..3 (src/WDPMCL.c:1184-1186), a colour
// sub-pass (oi,oj) relaxes the 3x3 tiles whose top-left corner is at
// row = oi-1 (mod 3), col = oj-1 (mod 3). For row-triple index m:
//      A_m = rows 3m   .. 3m+2   tiles of oi=1   (phase type q = 0)
//      B_m = rows 3m+1 .. 3m+3   tiles of oi=2   (q = 1)
//      C_m = rows 3m+2 .. 3m+4   tiles of oi=3   (q = 2)
// B_m needs A_m and A_{m+1} finished (all three oj), C_m needs B_m and B_{m+1},
// and the next iteration's A_m needs C_{m-1} and C_m. So the 3K phases of K
// iterations can run as a software pipeline down the rows: in step s, phase p
// works on triples  m_lo + NT*s - p*LAG + t  (t < NT) with LAG = NT+1, every
// phase of a step touching disjoint rows whose inputs were finished in earlier
// steps. A step is three block-wide sub-steps (oj = 1,2,3), each running all
// phases' tiles in parallel. Rows finished by the last phase are written back;
// new rows are prefetched PF steps ahead. Columns are not pipelined: validity
// shrinks by 1 column on the left and 2 on the right per sub-pass after the
// first, which the halos HL / HR absorb (tiles that do not fit entirely in the
// window are skipped; the cells they would have produced are never stored).
//
// Everything here is plain index arithmetic shared by the CUDA kernel
// (kernels.cuh) and the CPU schedule emulator used by the tests
// (tests/emul/mw_emul.cpp), so that the two cannot drift apart.
#pragma once

#ifdef __CUDACC__
#define WDPM_SCHED_HD __host__ __device__ __forceinline__
#else
#define WDPM_SCHED_HD inline
#endif

namespace wdpm {

constexpr int kPadLeft = 48;  // device columns left of padded column 0 (multiple of 12 and of 16 B / 4)
constexpr int kPadTop = 12;   // device rows above padded row 0 (>= 3 * max iterations per launch)
constexpr int kMaxItersPerLaunch = 4;

template <int W_, int NT_, int K_, int PF_>
struct MwCfg {
    static constexpr int W = W_;          // window columns staged in shared memory
    static constexpr int NT = NT_;        // row triples per phase per step
    static constexpr int K = K_;          // iterations per launch
    static constexpr int PF = PF_;        // prefetch distance in steps
    static constexpr int NPH = 3 * K;     // phases
    static constexpr int LAG = NT + 1;    // triple lag between consecutive phases
    static constexpr int HL = 12 * K;     // left halo columns  (needs >= 9K-1; multiple of 12 keeps 16 B alignment and mod-3 phase)
    static constexpr int HR_NEED = 18 * K - 2;  // right halo columns needed
    // tiles per row triple, the same for every tile-start column offset cofs = oj-1 in {0,1,2}:
    // tile c starts at window column 3c+cofs and must end inside the window for cofs = 2.
    static constexpr int NC = (W - 5) / 3 + 1;
    // Tiles are numbered row-group major; NCP is the numbering stride of a row group. Rounding it up
    // to a whole number of warps keeps every warp inside one row triple (stride-3 shared-memory
    // access is then conflict-free); it is done only when it idles fewer than 1 lane in 20.
    static constexpr int NCP = ((((NC + 31) / 32) * 32 - NC) * 20 < NC) ? ((NC + 31) / 32) * 32 : NC;
    // the oj=1 tessellation covers window columns [0, 3*NC): that is the width the halos eat into
    static constexpr int TWV = ((3 * NC - HL - HR_NEED) / 12) * 12;  // owned columns per strip
    static constexpr int TOP_TRIPLES = K;      // triples staged above the owned rows
    static constexpr int BOT_TRIPLES = 2 * K;  // triples staged below
    // Ring rows: from the oldest row whose write-back may still be reading shared memory (the
    // store group of the previous step is allowed to be in flight) to the newest prefetched row.
    static constexpr int NRING_MIN = 3 * NT * (PF + 2) + 3 * (NPH - 1) * LAG - 2;
#ifndef WDPM_NRING_DELTA
#define WDPM_NRING_DELTA 0 /* tests shrink the ring to prove the hazard checks fire */
#endif
    static constexpr int NRING = ((NRING_MIN + 2) / 3) * 3 + WDPM_NRING_DELTA;  // ring rows (multiple of 3)
    static constexpr int NSTAGE = PF + 1;      // load barriers
    static constexpr bool STRICT_ORDER = false;
    static_assert(W % 4 == 0, "window rows must be 16-byte multiples");
    static_assert(TWV > 0, "window too narrow for its halos");
    static_assert(K >= 1 && K <= kMaxItersPerLaunch, "iterations per launch");
    static_assert(HL <= kPadLeft, "left halo exceeds the device margin");
};

// Tiling of the WARP-AUTONOMOUS iteration kernel (kernels.cuh, k_fused_wa). Rows march exactly as above (same
// phases, lags, ring, loads and write-backs); what differs is who relaxes the tiles of a row triple.
// Every lane owns TWO column-adjacent tiles (6 columns) and keeps them, water and elevations, in a
// 3 x 8 register window for the three colour sub-passes of a step; the window slides one column per sub-pass:
// the entering column comes from the lane to the right by warp shuffle (that lane's leaving column). A
// warp is therefore self-contained: no shared-memory traffic and no barrier between sub-passes. The price is
// that a warp's last lane cannot slide (nobody to its right inside the warp): it recomputes the first
// columns of the next warp only to feed lane 30, and stores nothing. A warp spans 32*6 = 192 columns of which
// lanes 0..30 finish and store the columns [2, 188) relative to the warp's start; consecutive warps start
// WSTRIDE = 186 columns apart so that these ranges abut. The window's own edges lose 2 columns (left) and
// 4 (right) per phase exactly as the halo analysis above says (8 and 16 over the three phases of an iteration).
// RING_PF < PF ("deep prefetch"): the ring is sized for a prefetch distance of RING_PF steps although loads go
// out PF steps ahead. It works when the data-movement warp issues the loads of a step only AFTER the write-backs it
// has just issued have read their rows out of shared memory (cp.async.bulk.wait_group.read 0 between the two): the
// ring formula reserves 3*NT rows for a write-back that may still be reading while new rows land; with the strict
// order those rows can take the extra step of prefetch instead. (PF = 2 on the PF = 1 ring: the rows loaded at step s
// for step s+2 alias the rows stored at step s and nothing younger; the emulator's hazard counters check it.)
template <int KW_, int NT_, int PF_, int RING_PF_ = PF_>
struct WaCfg {
    static constexpr int KW = KW_;        // warps per row triple
    static constexpr int NT = NT_;        // row triples per phase per step
    static constexpr int K = 1;           // iterations per launch
    static constexpr int PF = PF_;        // prefetch distance in steps
    static constexpr int RING_PF = RING_PF_;
    static constexpr bool STRICT_ORDER = RING_PF_ < PF_;  // loads only after the step's write-backs were read out
    static_assert(RING_PF_ == PF_ || RING_PF_ == PF_ - 1, "the strict order buys exactly one step of prefetch");
    static constexpr int NPH = 3;
    static constexpr int LAG = NT + 1;
    static constexpr int CPL = 6;                    // columns (two tiles) per lane
    static constexpr int WSTRIDE = 31 * CPL;         // columns finished per warp = distance between warp starts
    static constexpr int W = ((KW * WSTRIDE + CPL + 2 + 3) / 4) * 4;  // last warp's lane 31 reads columns up to KW*WSTRIDE + 7
    static constexpr int HL = 12;                    // needs >= 8; multiple of 12 keeps 16 B alignment and the colour phase
    static constexpr int TWV = ((KW * WSTRIDE - 10 - HL) / 12) * 12;  // valid after three phases: [8, KW*WSTRIDE - 10)
    static constexpr int TOP_TRIPLES = 1;
    static constexpr int BOT_TRIPLES = 2;
    static constexpr int NRING_MIN = 3 * NT * (RING_PF + 2) + 3 * (NPH - 1) * LAG - 2;
    static constexpr int NRING = ((NRING_MIN + 2) / 3) * 3 + WDPM_NRING_DELTA;
    static constexpr int NSTAGE = PF + 1;
    static constexpr int NWARPS = NPH * NT * KW;     // compute warps
    static_assert(TWV > 0 && HL <= kPadLeft && W % 4 == 0, "window geometry");
};

// Per-CTA view of the schedule. All rows/cols are PADDED grid coordinates
// (row 0 / col 0 = the reference's halo ring); they may be negative or exceed the
// grid inside the device margins, which hold dem = nodata, water = 0.
template <typename CFG>
struct MwTile {
    int x0;    // padded column of window column 0   (= strip*TWV - HL, always = 0 mod 3)
    int m0;    // first owned triple: owned rows are [3*m0, 3*m1)
    int m1;
    int m_lo;  // first / last staged triple (A-type rows 3m..3m+2)
    int m_hi;
    int n_steps;

    WDPM_SCHED_HD void init(int strip, int chunk, int chunk_triples, int total_triples) {
        x0 = strip * CFG::TWV - CFG::HL;
        m0 = chunk * chunk_triples;
        m1 = m0 + chunk_triples;
        if (m1 > total_triples) m1 = total_triples;
        m_lo = m0 - CFG::TOP_TRIPLES;
        m_hi = m1 + CFG::BOT_TRIPLES - 1;
        const int span = m_hi - m_lo + 1 + (CFG::NPH - 1) * CFG::LAG;
        n_steps = (span + CFG::NT - 1) / CFG::NT;
    }
    // triple handled by phase p, slot t of step s
    WDPM_SCHED_HD int triple(int s, int p, int t) const { return m_lo + CFG::NT * s - p * CFG::LAG + t; }
    // may phase type q (0,1,2) run on triple m?  (all three rows staged)
    WDPM_SCHED_HD bool runnable(int m, int q) const { return m >= m_lo && m <= m_hi - (q > 0 ? 1 : 0); }
    // is triple m staged by a load (A-type rows)?
    WDPM_SCHED_HD bool staged(int m) const { return m >= m_lo && m <= m_hi; }
    WDPM_SCHED_HD int first_row() const { return 3 * m_lo; }
    WDPM_SCHED_HD int ring_slot(int row) const { return (row - 3 * m_lo) % CFG::NRING; }
    WDPM_SCHED_HD bool owns_row(int row) const { return row >= 3 * m0 && row < 3 * m1; }
    WDPM_SCHED_HD bool owns_col(int col) const { return col >= x0 + CFG::HL && col < x0 + CFG::HL + CFG::TWV; }
};

}  // namespace wdpm
