// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
//
// CPU emulation of the fused kernel's schedule (wdpm_b200/csrc/kernels.cuh,
// k_fused). It executes, CTA by CTA and step by step, exactly the loads, tile
// relaxations and write-backs the CUDA kernel issues - using the SAME index
// arithmetic (wdpm_b200/csrc/mw_schedule.h) and the SAME relax functions
// (wdpm_b200/csrc/relax.cuh, compiled for the host) - on a model of the
// shared-memory row ring. It exists so that halo widths, pipeline lags and ring
// sizing can be proven against the oracle on a machine without a GPU, and it
// checks the hazards a GPU run would only show as silent corruption:
//   * every ring slot carries the id of the row it holds; a tile that finds a
//     different row than it expects (slot reloaded too early, or not loaded yet)
//     is an error;
//   * a load into a slot whose write-back has not been retired by the
//     corresponding bulk wait is an error (loads are modelled as landing
//     immediately, stores as reading as late as the kernel allows).
// Nothing under wdpm_b200/ links against this file.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <vector>

#include "../../wdpm_b200/csrc/mw_schedule.h"
#include "../../wdpm_b200/csrc/relax.cuh"

using namespace wdpm;

namespace {

struct Layout {
    int R, C, pitch, nrows_dev;
    size_t at(int i, int j) const { return (size_t)(i + kPadTop) * pitch + (size_t)(j + kPadLeft); }
};

template <typename T>
struct Event { T w_outlet, w_centre; int valid; };

struct Errors {
    long long wrong_row = 0;      // tile touched a slot holding another row
    long long load_over_store = 0;  // load landed in a slot with an unretired store
    long long double_store = 0;
    long long unstored = 0;
};

template <typename T, int MODULE, typename CFG>
void run_cta(const Layout& L, const T* w_in, T* w_out, const T* dem, T nodata, int strip, int chunk,
             int chunk_triples, int total_triples, int drainrow, int draincol, Event<T>* events, Errors& err,
             std::vector<unsigned char>& stored_mask) {
    constexpr int W = CFG::W, NT = CFG::NT, NPH = CFG::NPH, NRING = CFG::NRING, PF = CFG::PF;
    MwTile<CFG> tile;
    tile.init(strip, chunk, chunk_triples, total_triples);
    std::vector<T> ring_w((size_t)NRING * W), ring_d((size_t)NRING * W);
    std::vector<int> slot_row(NRING, INT32_MIN);
    // store groups in flight: each is a list of slots
    std::deque<std::vector<int>> groups;
    std::vector<int> slot_pending(NRING, 0);
    const int col0 = tile.x0 + kPadLeft;

    auto issue_loads = [&](int s) {
        for (int t = 0; t < NT; t++) {
            const int m = tile.triple(s, 0, t);
            if (!tile.staged(m)) continue;
            for (int k = 0; k < 3; k++) {
                const int row = 3 * m + k;
                const int slot = tile.ring_slot(row);
                if (slot_pending[slot]) err.load_over_store++;
                const size_t src = (size_t)(row + kPadTop) * L.pitch + col0;
                std::memcpy(&ring_w[(size_t)slot * W], w_in + src, W * sizeof(T));
                std::memcpy(&ring_d[(size_t)slot * W], dem + src, W * sizeof(T));
                slot_row[slot] = row;
            }
        }
    };
    auto wait_read = [&](size_t allowed) {
        while (groups.size() > allowed) {
            for (int slot : groups.front()) slot_pending[slot]--;
            groups.pop_front();
        }
    };

    for (int s = 0; s < PF && s < tile.n_steps; s++) issue_loads(s);
    for (int s = 0; s < tile.n_steps; s++) {
        if (s + PF < tile.n_steps) issue_loads(s + PF);
        for (int cofs = 0; cofs < 3; cofs++) {
            constexpr int nc = CFG::NC;
            for (int item = 0; item < NPH * NT * nc; item++) {
                const int c = item % nc, pt = item / nc, t = pt % NT, ph = pt / NT, q = ph % 3;
                const int m = tile.triple(s, ph, t);
                if (!tile.runnable(m, q)) continue;
                const int row0 = 3 * m + q;
                int s0 = tile.ring_slot(row0);
                int s1 = s0 + 1; if (s1 == NRING) s1 = 0;
                int s2 = s1 + 1; if (s2 == NRING) s2 = 0;
                if (slot_row[s0] != row0 || slot_row[s1] != row0 + 1 || slot_row[s2] != row0 + 2) err.wrong_row++;
                const int j = 3 * c + cofs + 1;
                T* w0 = &ring_w[(size_t)s0 * W]; T* w1 = &ring_w[(size_t)s1 * W]; T* w2 = &ring_w[(size_t)s2 * W];
                const T* d0 = &ring_d[(size_t)s0 * W]; const T* d1 = &ring_d[(size_t)s1 * W]; const T* d2 = &ring_d[(size_t)s2 * W];
                if (MODULE == kDrain) {
                    const int crow = row0 + 1, ccol = tile.x0 + j;
                    const int orow = drainrow - crow, ocol = draincol - ccol;
                    if (orow >= -1 && orow <= 1 && ocol >= -1 && ocol <= 1) {
                        T evo[8], evc[8]; int pos[8];
                        const int n = relax_tile_near_outlets<T>(w0, w1, w2, d0, d1, d2, j, 1 << ((orow + 1) * 3 + (ocol + 1)), evo, evc, pos);
                        if (n && tile.owns_row(crow) && tile.owns_col(ccol)) {
                            Event<T>& e = events[(ph / 3) * 9 + q * 3 + cofs];
                            e.w_outlet = evo[0]; e.w_centre = evc[0]; e.valid = 1;
                        }
                        continue;
                    }
                }
                relax_tile<T, MODULE>(w0, w1, w2, d0, d1, d2, j);
            }
        }
        std::vector<int> grp;
        for (int t = 0; t < NT; t++) {
            const int m = tile.triple(s, NPH - 1, t);
            for (int k = 0; k < 3; k++) {
                const int row = 3 * m + 2 + k;
                if (!tile.owns_row(row)) continue;
                const int slot = tile.ring_slot(row);
                if (slot_row[slot] != row) err.wrong_row++;
                const size_t dst = (size_t)(row + kPadTop) * L.pitch + col0 + CFG::HL;
                std::memcpy(w_out + dst, &ring_w[(size_t)slot * W + CFG::HL], CFG::TWV * sizeof(T));
                for (int c = 0; c < CFG::TWV; c++) {
                    if (stored_mask[dst + c]) err.double_store++;
                    stored_mask[dst + c] = 1;
                }
                grp.push_back(slot);
                slot_pending[slot]++;
            }
        }
        if (!grp.empty()) groups.push_back(grp);
        wait_read(1);
    }
    wait_read(0);
}

template <typename T, int MODULE, typename CFG>
int run_launches(T* w_padded, const T* d_padded, int R, int C, T nodata, int n_launches, int chunk_triples,
                 int drainrow, int draincol, T* totaldrain, long long* err_out) {
    Layout L;
    L.R = R; L.C = C;
    const int n_strips = (C + 2 + CFG::TWV - 1) / CFG::TWV;
    const int total_triples = (R + 2 + 2) / 3;
    L.pitch = ((kPadLeft + n_strips * CFG::TWV + (CFG::W - CFG::TWV - CFG::HL) + 31) / 32) * 32;
    L.nrows_dev = kPadTop + 3 * (total_triples + 2 * CFG::K) + 3;
    if (chunk_triples <= 0) chunk_triples = total_triples;
    const int n_chunks = (total_triples + chunk_triples - 1) / chunk_triples;
    const size_t n = (size_t)L.pitch * L.nrows_dev;
    std::vector<T> dem(n, invalid_elevation<T>()), wa(n, T(0)), wb(n, T(0));  // masked elevations, as the solver stores them
    for (int i = 0; i < R + 2; i++)
        for (int j = 0; j < C + 2; j++) {
            dem[L.at(i, j)] = mask_elevation(d_padded[(size_t)i * (C + 2) + j], nodata);
            wa[L.at(i, j)] = w_padded[(size_t)i * (C + 2) + j];
        }
    Errors err;
    T* cur = wa.data();
    T* nxt = wb.data();
    T td = *totaldrain;
    for (int l = 0; l < n_launches; l++) {
        std::vector<Event<T>> events(9 * CFG::K, Event<T>{T(0), T(0), 0});
        std::vector<unsigned char> stored(n, 0);
        for (int chunk = 0; chunk < n_chunks; chunk++)
            for (int strip = 0; strip < n_strips; strip++)
                run_cta<T, MODULE, CFG>(L, cur, nxt, dem.data(), nodata, strip, chunk, chunk_triples, total_triples,
                                        drainrow, draincol, events.data(), err, stored);
        for (int i = 0; i < R + 2; i++)
            for (int j = 0; j < C + 2; j++)
                if (!stored[L.at(i, j)]) err.unstored++;
        for (auto& e : events)
            if (e.valid) { td = td + e.w_outlet; td = td + e.w_centre; }
        T* tmp = cur; cur = nxt; nxt = tmp;
    }
    // margins of the final buffer must still be zero
    long long margin_dirty = 0;
    {
        std::vector<unsigned char> inside(n, 0);
        for (int i = 0; i < R + 2; i++)
            for (int j = 0; j < C + 2; j++) inside[L.at(i, j)] = 1;
        for (size_t k = 0; k < n; k++)
            if (!inside[k] && cur[k] != T(0)) margin_dirty++;
    }
    for (int i = 0; i < R + 2; i++)
        for (int j = 0; j < C + 2; j++) w_padded[(size_t)i * (C + 2) + j] = cur[L.at(i, j)];
    *totaldrain = td;
    err_out[0] = err.wrong_row;
    err_out[1] = err.load_over_store;
    err_out[2] = err.double_store;
    err_out[3] = err.unstored;
    err_out[4] = margin_dirty;
    return 0;
}

// Warp-autonomous variant (k_fused_wa): the same row march; the tiles of a row triple are relaxed by KW warps
// of 32 lanes, each lane holding a 3 x 8 window that slides by warp shuffle. Modelled exactly: all warps of a
// row group read their windows before any of them writes back (the kernel's row-group mbarrier), lane 31
// stores nothing, shuffles take the right-hand lane's value (lane 31 keeps its own).
template <typename T, int MODULE, typename CFG, bool FAST, bool GUARD>
void run_cta_wa(const Layout& L, const T* w_in, T* w_out, const T* dem, int strip, int chunk,
                int chunk_triples, int total_triples, Errors& err, std::vector<unsigned char>& stored_mask, Event<T>* events = nullptr) {
    constexpr int W = CFG::W, NT = CFG::NT, NPH = CFG::NPH, NRING = CFG::NRING, PF = CFG::PF, KW = CFG::KW;
    MwTile<CFG> tile;
    tile.init(strip, chunk, chunk_triples, total_triples);
    std::vector<T> ring_w((size_t)NRING * W), ring_d((size_t)NRING * W);
    std::vector<int> slot_row(NRING, INT32_MIN);
    std::deque<std::vector<int>> groups;
    std::vector<int> slot_pending(NRING, 0);
    const int col0 = tile.x0 + kPadLeft;

    auto issue_loads = [&](int s) {
        for (int t = 0; t < NT; t++) {
            const int m = tile.triple(s, 0, t);
            if (!tile.staged(m)) continue;
            for (int k = 0; k < 3; k++) {
                const int row = 3 * m + k;
                const int slot = tile.ring_slot(row);
                if (slot_pending[slot]) err.load_over_store++;
                const size_t src = (size_t)(row + kPadTop) * L.pitch + col0;
                std::memcpy(&ring_w[(size_t)slot * W], w_in + src, W * sizeof(T));
                std::memcpy(&ring_d[(size_t)slot * W], dem + src, W * sizeof(T));
                slot_row[slot] = row;
            }
        }
    };
    auto wait_read = [&](size_t allowed) {
        while (groups.size() > allowed) {
            for (int slot : groups.front()) slot_pending[slot]--;
            groups.pop_front();
        }
    };
    struct Lane { T wt[3][8], dd[3][8]; };
    std::vector<Lane> lanes((size_t)KW * 32);

    for (int s = 0; s < PF && s < tile.n_steps; s++) issue_loads(s);
    for (int s = 0; s < tile.n_steps; s++) {
        if (CFG::STRICT_ORDER) wait_read(0);  // deep prefetch: the write-backs issued at the end of the step before were read out first
        if (s + PF < tile.n_steps) issue_loads(s + PF);
        for (int grp = 0; grp < NPH * NT; grp++) {
            const int t = grp % NT, ph = grp / NT;
            const int m = tile.triple(s, ph, t);
            if (!tile.runnable(m, ph)) continue;
            const int row0 = 3 * m + ph;
            int sl[3];
            sl[0] = tile.ring_slot(row0);
            sl[1] = sl[0] + 1 == NRING ? 0 : sl[0] + 1;
            sl[2] = sl[1] + 1 == NRING ? 0 : sl[1] + 1;
            for (int r = 0; r < 3; r++)
                if (slot_row[sl[r]] != row0 + r) err.wrong_row++;
            for (int kw = 0; kw < KW; kw++)
                for (int lane = 0; lane < 32; lane++) {
                    Lane& ln = lanes[(size_t)kw * 32 + lane];
                    const int cb = CFG::WSTRIDE * kw + CFG::CPL * lane;
                    if (cb + 8 > W) { err.wrong_row++; continue; }
                    for (int r = 0; r < 3; r++) {
                        for (int c = 0; c < 6; c++) ln.wt[r][c] = ring_w[(size_t)sl[r] * W + cb + c];
                        for (int c = 0; c < 8; c++) ln.dd[r][c] = ring_d[(size_t)sl[r] * W + cb + c];
                        ln.wt[r][6] = ln.wt[r][7] = T(0);
                    }
                }
            for (int kw = 0; kw < KW; kw++) {
                Lane* wl = &lanes[(size_t)kw * 32];
                // Drain: a warp with an outlet mark in any lane's window takes the outlet path (the kernel's __any_sync)
                bool near = false;
                if (MODULE == kDrain)
                    for (int lane = 0; lane < 32; lane++)
                        for (int r = 0; r < 3; r++)
                            for (int c = 0; c < 8; c++) near = near || is_outlet(wl[lane].dd[r][c]);
                const int crow = row0 + 1;
                auto sub = [&](auto c_tag) {
                    constexpr int CC = decltype(c_tag)::value;
                    for (int lane = 0; lane < 32; lane++) {
                        if (MODULE == kDrain && near) {
                            const int cb = CFG::WSTRIDE * kw + CFG::CPL * lane;
                            wa_relax_pair_outlets<T, CC, FAST>(wl[lane].wt, wl[lane].dd, [&](int tl, int a, int b, T wo, T wc) {
                                (void)a; (void)b;
                                const int ccol = tile.x0 + cb + CC + 3 * tl + 1;
                                if (lane < 31 && tile.owns_row(crow) && tile.owns_col(ccol) && events) {
                                    Event<T>& e = events[ph * 3 + CC];
                                    if (e.valid) err.double_store++;  // one contact per outlet and sub-pass
                                    e.w_outlet = wo; e.w_centre = wc; e.valid = 1;
                                }
                            });
                        } else {
                            wa_relax_pair<T, MODULE, CC, FAST, GUARD>(wl[lane].wt, wl[lane].dd);
                        }
                    }
                };
                sub(std::integral_constant<int, 0>{});
                for (int lane = 0; lane < 32; lane++)
                    for (int r = 0; r < 3; r++) wl[lane].wt[r][6] = wl[lane < 31 ? lane + 1 : lane].wt[r][0];
                sub(std::integral_constant<int, 1>{});
                for (int lane = 0; lane < 32; lane++)
                    for (int r = 0; r < 3; r++) wl[lane].wt[r][7] = wl[lane < 31 ? lane + 1 : lane].wt[r][1];
                sub(std::integral_constant<int, 2>{});
            }
            for (int kw = 0; kw < KW; kw++)
                for (int lane = 0; lane < 31; lane++) {
                    const Lane& ln = lanes[(size_t)kw * 32 + lane];
                    const int cb = CFG::WSTRIDE * kw + CFG::CPL * lane;
                    for (int r = 0; r < 3; r++)
                        for (int c = 2; c < 8; c++) ring_w[(size_t)sl[r] * W + cb + c] = ln.wt[r][c];
                }
        }
        std::vector<int> grp;
        for (int t = 0; t < NT; t++) {
            const int m = tile.triple(s, NPH - 1, t);
            for (int k = 0; k < 3; k++) {
                const int row = 3 * m + 2 + k;
                if (!tile.owns_row(row)) continue;
                const int slot = tile.ring_slot(row);
                if (slot_row[slot] != row) err.wrong_row++;
                const size_t dst = (size_t)(row + kPadTop) * L.pitch + col0 + CFG::HL;
                std::memcpy(w_out + dst, &ring_w[(size_t)slot * W + CFG::HL], CFG::TWV * sizeof(T));
                for (int c = 0; c < CFG::TWV; c++) {
                    if (stored_mask[dst + c]) err.double_store++;
                    stored_mask[dst + c] = 1;
                }
                grp.push_back(slot);
                slot_pending[slot]++;
            }
        }
        if (!grp.empty()) groups.push_back(grp);
        wait_read(1);
    }
    wait_read(0);
}

// The STAGGERED order of k_fused_wa (kOptStagger, NT = 2): half steps; slot-0 groups begin step h/2 at even h and
// end it at odd h, slot-1 groups half a step later; rows are written home half a step after they are finished,
// slot 0 and slot 1 in store groups of their own. Windows live in "registers" (Lane) between the halves, so a
// ring slot that is reloaded or overwritten in between shows up as a wrong row at write-back.
template <typename T, int MODULE, typename CFG, bool FAST, bool GUARD>
void run_cta_wa_stag(const Layout& L, const T* w_in, T* w_out, const T* dem, int strip, int chunk,
                     int chunk_triples, int total_triples, Errors& err, std::vector<unsigned char>& stored_mask) {
    constexpr int W = CFG::W, NT = CFG::NT, NPH = CFG::NPH, NRING = CFG::NRING, PF = CFG::PF, KW = CFG::KW;
    static_assert(NT == 2, "staggered schedule");
    MwTile<CFG> tile;
    tile.init(strip, chunk, chunk_triples, total_triples);
    std::vector<T> ring_w((size_t)NRING * W), ring_d((size_t)NRING * W);
    std::vector<int> slot_row(NRING, INT32_MIN);
    std::deque<std::vector<int>> groups;
    std::vector<int> slot_pending(NRING, 0);
    const int col0 = tile.x0 + kPadLeft;
    const int n = tile.n_steps;

    // everything inside a half step runs concurrently on the GPU: two parties (row groups, the copy engine) touching
    // the same ring slot in the same half step is a race
    std::vector<int> claim_h(NRING, -2), claim_who(NRING, -1);
    int cur_h = -1;
    auto claim = [&](int slot, int who) {
        if (claim_h[slot] == cur_h && claim_who[slot] != who) {
            err.wrong_row++;
            if (getenv("WDPM_EMUL_DEBUG")) fprintf(stderr, "claim conflict: half step %d slot %d row %d: %d vs %d\n", cur_h, slot, slot_row[slot], claim_who[slot], who);
        }
        claim_h[slot] = cur_h;
        claim_who[slot] = who;
    };

    auto issue_loads = [&](int s) {
        for (int t = 0; t < NT; t++) {
            const int m = tile.triple(s, 0, t);
            if (!tile.staged(m)) continue;
            for (int k = 0; k < 3; k++) {
                const int row = 3 * m + k;
                const int slot = tile.ring_slot(row);
                if (slot_pending[slot]) err.load_over_store++;
                claim(slot, 100);
                const size_t src = (size_t)(row + kPadTop) * L.pitch + col0;
                std::memcpy(&ring_w[(size_t)slot * W], w_in + src, W * sizeof(T));
                std::memcpy(&ring_d[(size_t)slot * W], dem + src, W * sizeof(T));
                slot_row[slot] = row;
            }
        }
    };
    auto issue_stores = [&](int s, int t) {
        std::vector<int> grp;
        const int m = tile.triple(s, NPH - 1, t);
        for (int k = 0; k < 3; k++) {
            const int row = 3 * m + 2 + k;
            if (!tile.owns_row(row)) continue;
            const int slot = tile.ring_slot(row);
            if (slot_row[slot] != row) err.wrong_row++;
            claim(slot, 101);
            const size_t dst = (size_t)(row + kPadTop) * L.pitch + col0 + CFG::HL;
            std::memcpy(w_out + dst, &ring_w[(size_t)slot * W + CFG::HL], CFG::TWV * sizeof(T));
            for (int c = 0; c < CFG::TWV; c++) {
                if (stored_mask[dst + c]) err.double_store++;
                stored_mask[dst + c] = 1;
            }
            grp.push_back(slot);
            slot_pending[slot]++;
        }
        if (!grp.empty()) groups.push_back(grp);
    };
    auto wait_read = [&](size_t allowed) {
        while (groups.size() > allowed) {
            for (int slot : groups.front()) slot_pending[slot]--;
            groups.pop_front();
        }
    };
    struct Lane { T wt[3][8], dd[3][8]; };
    struct Group { std::vector<Lane> lanes; bool run = false; int sl[3] = {0, 0, 0}; int row0 = 0; };
    std::vector<Group> gs(NPH * NT);
    for (auto& g : gs) g.lanes.resize((size_t)KW * 32);
    auto begin_step = [&](Group& g, int ph, int t, int s) {
        const int m = tile.triple(s, ph, t);
        g.run = tile.runnable(m, ph);
        if (!g.run) return;
        g.row0 = 3 * m + ph;
        g.sl[0] = tile.ring_slot(g.row0);
        g.sl[1] = g.sl[0] + 1 == NRING ? 0 : g.sl[0] + 1;
        g.sl[2] = g.sl[1] + 1 == NRING ? 0 : g.sl[1] + 1;
        for (int r = 0; r < 3; r++) {
            claim(g.sl[r], ph * NT + t);
            if (slot_pending[g.sl[r]]) err.load_over_store++;  // a row whose write-back is still reading it
        }
        for (int r = 0; r < 3; r++)
            if (slot_row[g.sl[r]] != g.row0 + r) err.wrong_row++;
        for (int kw = 0; kw < KW; kw++)
            for (int lane = 0; lane < 32; lane++) {
                Lane& ln = g.lanes[(size_t)kw * 32 + lane];
                const int cb = CFG::WSTRIDE * kw + CFG::CPL * lane;
                for (int r = 0; r < 3; r++) {
                    for (int c = 0; c < 6; c++) ln.wt[r][c] = ring_w[(size_t)g.sl[r] * W + cb + c];
                    for (int c = 0; c < 8; c++) ln.dd[r][c] = ring_d[(size_t)g.sl[r] * W + cb + c];
                    ln.wt[r][6] = ln.wt[r][7] = T(0);
                }
            }
        for (int kw = 0; kw < KW; kw++) {
            Lane* wl = &g.lanes[(size_t)kw * 32];
            for (int lane = 0; lane < 32; lane++) wa_relax_pair<T, MODULE, 0, FAST, GUARD>(wl[lane].wt, wl[lane].dd);
            for (int lane = 0; lane < 32; lane++)
                for (int r = 0; r < 3; r++) wl[lane].wt[r][6] = wl[lane < 31 ? lane + 1 : lane].wt[r][0];
            for (int lane = 0; lane < 32; lane++) wa_relax_pair<T, MODULE, 1, FAST, GUARD, 1>(wl[lane].wt, wl[lane].dd);
        }
    };
    auto end_step = [&](Group& g, int who) {
        if (!g.run) return;
        for (int r = 0; r < 3; r++) claim(g.sl[r], who);
        for (int kw = 0; kw < KW; kw++) {
            Lane* wl = &g.lanes[(size_t)kw * 32];
            for (int lane = 0; lane < 32; lane++) wa_relax_pair<T, MODULE, 1, FAST, GUARD, 2>(wl[lane].wt, wl[lane].dd);
            for (int lane = 0; lane < 32; lane++)
                for (int r = 0; r < 3; r++) wl[lane].wt[r][7] = wl[lane < 31 ? lane + 1 : lane].wt[r][1];
            for (int lane = 0; lane < 32; lane++) wa_relax_pair<T, MODULE, 2, FAST, GUARD>(wl[lane].wt, wl[lane].dd);
        }
        for (int r = 0; r < 3; r++)
            if (slot_row[g.sl[r]] != g.row0 + r) err.wrong_row++;  // the slot changed hands while the window was in registers
        for (int kw = 0; kw < KW; kw++)
            for (int lane = 0; lane < 31; lane++) {
                const Lane& ln = g.lanes[(size_t)kw * 32 + lane];
                const int cb = CFG::WSTRIDE * kw + CFG::CPL * lane;
                for (int r = 0; r < 3; r++)
                    for (int c = 2; c < 8; c++) ring_w[(size_t)g.sl[r] * W + cb + c] = ln.wt[r][c];
            }
    };

    for (int s = 0; s < PF && s < n; s++) issue_loads(s);
    for (int h = 0; h <= 2 * n; h++) {
        cur_h = h;
        {   // data-movement warp, before the compute of half step h
            const int s = h >> 1;
            if ((h & 1) == 0) {
                if (s > 0) { issue_stores(s - 1, 0); wait_read(1); }
                if (s + PF < n) issue_loads(s + PF);
            } else if (s > 0) {
                issue_stores(s - 1, 1);
            }
        }
        for (int grp = 0; grp < NPH * NT; grp++) {
            const int t = grp % NT, ph = grp / NT;
            const int s = (h - t) >> 1;
            if (((h - t) & 1) == 0) {
                if (h >= t && s < n) begin_step(gs[grp], ph, t, s);
            } else {
                if (h > t && s < n) end_step(gs[grp], grp);
            }
        }
    }
    issue_stores(n - 1, 1);
    wait_read(0);
}

template <typename T, int MODULE, typename CFG, bool FAST, bool GUARD>
int run_launches_wa(T* w_padded, const T* d_padded, int R, int C, T nodata, int n_launches, int chunk_triples, long long* err_out, bool stag,
                    int drainrow = -10, int draincol = -10, T* totaldrain = nullptr) {
    Layout L;
    L.R = R; L.C = C;
    const int n_strips = (C + 2 + CFG::TWV - 1) / CFG::TWV;
    const int total_triples = (R + 2 + 2) / 3;
    L.pitch = ((kPadLeft + n_strips * CFG::TWV + (CFG::W - CFG::TWV - CFG::HL) + 31) / 32) * 32;
    L.nrows_dev = kPadTop + 3 * (total_triples + 2 * kMaxItersPerLaunch) + 3;
    if (chunk_triples <= 0) chunk_triples = total_triples;
    const int n_chunks = (total_triples + chunk_triples - 1) / chunk_triples;
    const size_t n = (size_t)L.pitch * L.nrows_dev;
    std::vector<T> dem(n, invalid_elevation<T>()), wa(n, T(0)), wb(n, T(0));
    for (int i = 0; i < R + 2; i++)
        for (int j = 0; j < C + 2; j++) {
            dem[L.at(i, j)] = mask_elevation(d_padded[(size_t)i * (C + 2) + j], nodata);
            wa[L.at(i, j)] = w_padded[(size_t)i * (C + 2) + j];
        }
    if (MODULE == kDrain && drainrow >= 0 && drainrow <= R + 1 && draincol >= 0 && draincol <= C + 1)
        dem[L.at(drainrow, draincol)] = outlet_mark<T>();  // as the solver marks an outlet (relax.cuh)
    Errors err;
    T* cur = wa.data();
    T* nxt = wb.data();
    T td = totaldrain ? *totaldrain : T(0);
    for (int l = 0; l < n_launches; l++) {
        std::vector<Event<T>> events(9, Event<T>{T(0), T(0), 0});
        std::vector<unsigned char> stored(n, 0);
        for (int chunk = 0; chunk < n_chunks; chunk++)
            for (int strip = 0; strip < n_strips; strip++)
                if constexpr (CFG::NT == 2) {
                    if (stag) run_cta_wa_stag<T, MODULE, CFG, FAST, GUARD>(L, cur, nxt, dem.data(), strip, chunk, chunk_triples, total_triples, err, stored);
                    else run_cta_wa<T, MODULE, CFG, FAST, GUARD>(L, cur, nxt, dem.data(), strip, chunk, chunk_triples, total_triples, err, stored, events.data());
                } else {
                    run_cta_wa<T, MODULE, CFG, FAST, GUARD>(L, cur, nxt, dem.data(), strip, chunk, chunk_triples, total_triples, err, stored, events.data());
                }
        for (int i = 0; i < R + 2; i++)
            for (int j = 0; j < C + 2; j++)
                if (!stored[L.at(i, j)]) err.unstored++;
        for (auto& e : events)
            if (e.valid) { td = td + e.w_outlet; td = td + e.w_centre; }
        T* tmp = cur; cur = nxt; nxt = tmp;
    }
    long long margin_dirty = 0;
    {
        std::vector<unsigned char> inside(n, 0);
        for (int i = 0; i < R + 2; i++)
            for (int j = 0; j < C + 2; j++) inside[L.at(i, j)] = 1;
        for (size_t k = 0; k < n; k++)
            if (!inside[k] && cur[k] != T(0)) margin_dirty++;
    }
    for (int i = 0; i < R + 2; i++)
        for (int j = 0; j < C + 2; j++) w_padded[(size_t)i * (C + 2) + j] = cur[L.at(i, j)];
    if (totaldrain) *totaldrain = td;
    err_out[0] = err.wrong_row;
    err_out[1] = err.load_over_store;
    err_out[2] = err.double_store;
    err_out[3] = err.unstored;
    err_out[4] = margin_dirty;
    return 0;
}

// mode: bit 0 = fast Add step, bit 1 = no activity guard (Add on a clean grid only), bit 2 = staggered schedule (NT = 2)
template <typename T, typename CFG>
int dispatch_wa(int module, int mode, T* w, const T* d, int R, int C, T nodata, int n, int ct, long long* e, int dr = -10, int dc = -10, T* td = nullptr) {
    const bool stag = (mode & 4) != 0;
    if (stag && CFG::NT != 2) return -3;
    mode &= 3;
    if (module == kAdd && mode == 3) return run_launches_wa<T, kAdd, CFG, true, false>(w, d, R, C, nodata, n, ct, e, stag);
    if (module == kAdd && mode == 1) return run_launches_wa<T, kAdd, CFG, true, true>(w, d, R, C, nodata, n, ct, e, stag);
    if (module == kAdd) return run_launches_wa<T, kAdd, CFG, false, true>(w, d, R, C, nodata, n, ct, e, stag);
    if (module == kSubtract) return run_launches_wa<T, kSubtract, CFG, false, true>(w, d, R, C, nodata, n, ct, e, stag);
    if (module == kDrain && !stag) {
        if (mode & 1) return run_launches_wa<T, kDrain, CFG, true, true>(w, d, R, C, nodata, n, ct, e, false, dr, dc, td);
        return run_launches_wa<T, kDrain, CFG, false, true>(w, d, R, C, nodata, n, ct, e, false, dr, dc, td);
    }
    return -1;
}

template <typename T>
int dispatch_wa_cfg(int cfg, int module, int mode, T* w, const T* d, int R, int C, T nodata, int n, int ct, long long* e, int dr = -10, int dc = -10, T* td = nullptr) {
    switch (cfg) {
        case 0: return dispatch_wa<T, WaCfg<1, 1, 1>>(module, mode, w, d, R, C, nodata, n, ct, e, dr, dc, td);
        case 1: return dispatch_wa<T, WaCfg<2, 2, 1>>(module, mode, w, d, R, C, nodata, n, ct, e, dr, dc, td);
        case 2: return dispatch_wa<T, WaCfg<2, 1, 2>>(module, mode, w, d, R, C, nodata, n, ct, e, dr, dc, td);
        case 3: return dispatch_wa<T, WaCfg<3, 1, 1>>(module, mode, w, d, R, C, nodata, n, ct, e, dr, dc, td);
        case 4: return dispatch_wa<T, WaCfg<1, 2, 1>>(module, mode, w, d, R, C, nodata, n, ct, e, dr, dc, td);
        case 5: return dispatch_wa<T, WaCfg<2, 2, 2, 1>>(module, mode, w, d, R, C, nodata, n, ct, e, dr, dc, td);
        case 6: return dispatch_wa<T, WaCfg<1, 1, 2, 1>>(module, mode, w, d, R, C, nodata, n, ct, e, dr, dc, td);
    }
    return -2;
}

template <typename T, typename CFG>
int dispatch_module(int module, T* w, const T* d, int R, int C, T nodata, int n, int ct, int dr, int dc, T* td, long long* e) {
    switch (module) {
        case kAdd: return run_launches<T, kAdd, CFG>(w, d, R, C, nodata, n, ct, dr, dc, td, e);
        case kSubtract: return run_launches<T, kSubtract, CFG>(w, d, R, C, nodata, n, ct, dr, dc, td, e);
        case kDrain: return run_launches<T, kDrain, CFG>(w, d, R, C, nodata, n, ct, dr, dc, td, e);
    }
    return -1;
}

template <typename T>
int dispatch_cfg(int cfg, int module, T* w, const T* d, int R, int C, T nodata, int n, int ct, int dr, int dc, T* td, long long* e) {
    switch (cfg) {
        case 0: return dispatch_module<T, MwCfg<512, 1, 1, 2>>(module, w, d, R, C, nodata, n, ct, dr, dc, td, e);
        case 1: return dispatch_module<T, MwCfg<64, 1, 1, 1>>(module, w, d, R, C, nodata, n, ct, dr, dc, td, e);
        case 2: return dispatch_module<T, MwCfg<64, 1, 1, 2>>(module, w, d, R, C, nodata, n, ct, dr, dc, td, e);
        case 3: return dispatch_module<T, MwCfg<96, 2, 1, 1>>(module, w, d, R, C, nodata, n, ct, dr, dc, td, e);
        case 4: return dispatch_module<T, MwCfg<128, 1, 2, 2>>(module, w, d, R, C, nodata, n, ct, dr, dc, td, e);
        case 5: return dispatch_module<T, MwCfg<128, 2, 2, 1>>(module, w, d, R, C, nodata, n, ct, dr, dc, td, e);
        case 6: return dispatch_module<T, MwCfg<256, 1, 4, 3>>(module, w, d, R, C, nodata, n, ct, dr, dc, td, e);
        case 7: return dispatch_module<T, MwCfg<384, 2, 1, 2>>(module, w, d, R, C, nodata, n, ct, dr, dc, td, e);
    }
    return -2;
}

template <typename CFG>
static int cfg_info(int* W, int* TWV, int* K, int* NRING) { *W = CFG::W; *TWV = CFG::TWV; *K = CFG::K; *NRING = CFG::NRING; return 0; }
}  // namespace

extern "C" {
int mw_emul_cfg_info(int cfg, int* W, int* TWV, int* K, int* NRING) {
    switch (cfg) {
        case 0: return cfg_info<MwCfg<512, 1, 1, 2>>(W, TWV, K, NRING);
        case 1: return cfg_info<MwCfg<64, 1, 1, 1>>(W, TWV, K, NRING);
        case 2: return cfg_info<MwCfg<64, 1, 1, 2>>(W, TWV, K, NRING);
        case 3: return cfg_info<MwCfg<96, 2, 1, 1>>(W, TWV, K, NRING);
        case 4: return cfg_info<MwCfg<128, 1, 2, 2>>(W, TWV, K, NRING);
        case 5: return cfg_info<MwCfg<128, 2, 2, 1>>(W, TWV, K, NRING);
        case 6: return cfg_info<MwCfg<256, 1, 4, 3>>(W, TWV, K, NRING);
        case 7: return cfg_info<MwCfg<384, 2, 1, 2>>(W, TWV, K, NRING);
    }
    return -1;
}
int mw_emul_run_f64(int cfg, int module, double* w, const double* d, int R, int C, double nodata, int n_launches,
                    int chunk_triples, int drainrow, int draincol, double* totaldrain, long long* errors5) {
    return dispatch_cfg<double>(cfg, module, w, d, R, C, nodata, n_launches, chunk_triples, drainrow, draincol, totaldrain, errors5);
}
int wa_emul_cfg_info(int cfg, int* W, int* TWV) {
    switch (cfg) {
        case 0: *W = WaCfg<1, 1, 1>::W; *TWV = WaCfg<1, 1, 1>::TWV; return 0;
        case 1: *W = WaCfg<2, 2, 1>::W; *TWV = WaCfg<2, 2, 1>::TWV; return 0;
        case 2: *W = WaCfg<2, 1, 2>::W; *TWV = WaCfg<2, 1, 2>::TWV; return 0;
        case 3: *W = WaCfg<3, 1, 1>::W; *TWV = WaCfg<3, 1, 1>::TWV; return 0;
        case 4: *W = WaCfg<1, 2, 1>::W; *TWV = WaCfg<1, 2, 1>::TWV; return 0;
        case 5: *W = WaCfg<2, 2, 2, 1>::W; *TWV = WaCfg<2, 2, 2, 1>::TWV; return 0;
        case 6: *W = WaCfg<1, 1, 2, 1>::W; *TWV = WaCfg<1, 1, 2, 1>::TWV; return 0;
    }
    return -1;
}
int wa_emul_run_f64(int cfg, int module, int mode, double* w, const double* d, int R, int C, double nodata, int n_launches,
                    int chunk_triples, long long* errors5) {
    return dispatch_wa_cfg<double>(cfg, module, mode, w, d, R, C, nodata, n_launches, chunk_triples, errors5);
}
int wa_emul_run_f32(int cfg, int module, int mode, float* w, const float* d, int R, int C, float nodata, int n_launches,
                    int chunk_triples, long long* errors5) {
    return dispatch_wa_cfg<float>(cfg, module, mode, w, d, R, C, nodata, n_launches, chunk_triples, errors5);
}
int wa_emul_drain_f64(int cfg, int mode, double* w, const double* d, int R, int C, double nodata, int n_launches, int chunk_triples,
                      int drainrow, int draincol, double* totaldrain, long long* errors5) {
    return dispatch_wa_cfg<double>(cfg, kDrain, mode, w, d, R, C, nodata, n_launches, chunk_triples, errors5, drainrow, draincol, totaldrain);
}
int wa_emul_drain_f32(int cfg, int mode, float* w, const float* d, int R, int C, float nodata, int n_launches, int chunk_triples,
                      int drainrow, int draincol, float* totaldrain, long long* errors5) {
    return dispatch_wa_cfg<float>(cfg, kDrain, mode, w, d, R, C, nodata, n_launches, chunk_triples, errors5, drainrow, draincol, totaldrain);
}
int mw_emul_run_f32(int cfg, int module, float* w, const float* d, int R, int C, float nodata, int n_launches,
                    int chunk_triples, int drainrow, int draincol, float* totaldrain, long long* errors5) {
    return dispatch_cfg<float>(cfg, module, w, d, R, C, nodata, n_launches, chunk_triples, drainrow, draincol, totaldrain, errors5);
}
}
