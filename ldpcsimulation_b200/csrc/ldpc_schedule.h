// ldpc_schedule.h -- host-side, once per decoder: a bank-conflict-free order for the check-node
// gathers of a regular code.
//
// In the check phase lane l of warp w owns row 32w+l and, at step t, touches the message of one of
// its row's edges.  A message lives at word  s*N + col(i)  of shared memory, i.e. in bank
// col(i) mod 32 (N is a multiple of 32).  The alist's natural order makes a warp's 32 accesses of a
// step fall ~3.2 per bank (profiles/r1b: half of all shared-memory wavefronts are conflict replays).
// Min-sum's row update is order independent (min, second min and a sign XOR are exact and
// commutative), so the order in which a row visits its edges is free, and so is the storage column
// of a variable as long as the variable phase still walks columns at stride 1.  Hence:
//   1. choose col(): a bijection variable -> storage column such that, for every warp, its
//      32*dc edges fall exactly dc per bank   (balanced partition, found by local search);
//   2. per warp, the lane x bank multigraph is then dc-regular and bipartite, so by Koenig's theorem
//      it splits into dc perfect matchings: step t = matching t, 32 lanes -> 32 distinct banks.
// The result is a table of dc steps per row with no idle slots and no conflicts; decisions are
// bit-identical to the natural order.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <numeric>
#include <random>
#include <vector>

namespace ldpc {

struct RowSchedule {
    bool ok = false;
    std::vector<int> col;        // [N]   storage column of variable i
    std::vector<int> var_of_col; // [N]   inverse
    std::vector<int> order;      // [M*dc] for row j, step t: the slot k of its mlist row visited at step t
};

// mlist: [M*dc] 0-based variable of slot k of row j (regular: every row has dc entries).
inline RowSchedule build_row_schedule(int N, int M, int dc, const std::vector<int> &mlist, unsigned seed = 12345)
{
    RowSchedule rs;
    const int B = 32;
    if (N % B || dc < 1) return rs;
    const int W = (M + B - 1) / B, per_bank = N / B;              // the last warp may own fewer than 32 rows
    // ---- 1. balanced bank assignment -----------------------------------------------------------
    // edges of variable i as a list of warps (with multiplicity)
    std::vector<std::vector<int>> warps_of(N);
    for (int j = 0; j < M; j++) for (int k = 0; k < dc; k++) warps_of[mlist[(size_t)j * dc + k]].push_back(j / B);
    std::vector<int> bank(N);
    for (int i = 0; i < N; i++) bank[i] = i % B;
    std::vector<int> load((size_t)W * B, 0);
    for (int i = 0; i < N; i++) for (int w : warps_of[i]) load[(size_t)w * B + bank[i]]++;
    // a warp's lane x bank multigraph splits into dc conflict-free steps iff no bank receives more than dc of its edges
    // (Koenig: colours = maximum degree); for a full warp (32 dc edges) that means exactly dc per bank
    auto sq = [&](int v) { const long long d = v > dc ? v - dc : 0; return d * d; };
    long long cost = 0;
    for (int v : load) cost += sq(v);
    std::mt19937 rng(seed);
    // swapping the banks of two variables keeps every bank at N/32 variables
    auto delta_swap = [&](int a, int b) {
        const int ba = bank[a], bb = bank[b];
        long long d = 0;
        // apply tentatively on a small scratch map of (warp,bank) -> delta
        int tw[64], tb[64], tv[64], n = 0;
        auto add = [&](int w, int bk, int v) {
            for (int q = 0; q < n; q++) if (tw[q] == w && tb[q] == bk) { tv[q] += v; return; }
            tw[n] = w; tb[n] = bk; tv[n] = v; n++;
        };
        for (int w : warps_of[a]) { add(w, ba, -1); add(w, bb, +1); }
        for (int w : warps_of[b]) { add(w, bb, -1); add(w, ba, +1); }
        for (int q = 0; q < n; q++) { const int cur = load[(size_t)tw[q] * B + tb[q]]; d += sq(cur + tv[q]) - sq(cur); }
        return d;
    };
    auto do_swap = [&](int a, int b) {
        const int ba = bank[a], bb = bank[b];
        for (int w : warps_of[a]) { load[(size_t)w * B + ba]--; load[(size_t)w * B + bb]++; }
        for (int w : warps_of[b]) { load[(size_t)w * B + bb]--; load[(size_t)w * B + ba]++; }
        bank[a] = bb; bank[b] = ba;
    };
    for (int i = 0; i < N; i++) if (warps_of[i].size() > 15) return rs;   // scratch map above holds 4*dv entries
    const long long max_tries = 40LL * 1000 * 1000;
    for (long long it = 0; it < max_tries && cost > 0; it++) {
        const int a = (int)(rng() % N), b = (int)(rng() % N);
        if (bank[a] == bank[b]) continue;
        const long long d = delta_swap(a, b);
        // plain descent with sideways moves; the landscape is benign (few constraints per variable)
        if (d < 0 || (d == 0 && (rng() & 3) == 0)) { do_swap(a, b); cost += d; }
    }
    if (cost != 0) return rs;
    // storage column = bank + 32 * (rank of the variable inside its bank)
    rs.col.assign(N, -1); rs.var_of_col.assign(N, -1);
    std::vector<int> fill(B, 0);
    for (int i = 0; i < N; i++) {
        const int c = bank[i] + B * fill[bank[i]]++;
        rs.col[i] = c; rs.var_of_col[c] = i;
    }
    for (int b = 0; b < B; b++) if (fill[b] != per_bank) return rs;
    // ---- 2. per warp: dc-regular bipartite multigraph (lane x bank) -> dc perfect matchings ----
    rs.order.assign((size_t)M * dc, -1);
    std::vector<int> lane_col((size_t)B * dc), bank_col((size_t)B * dc);   // [node][colour] -> edge id or -1
    for (int w = 0; w < W; w++) {
        const int rows = std::min(B, M - w * B), nE = rows * dc;
        std::vector<int> eu(nE), ev(nE), ecol(nE, -1);                      // edge e = lane*dc + k
        for (int l = 0; l < rows; l++) for (int k = 0; k < dc; k++) { eu[l * dc + k] = l; ev[l * dc + k] = bank[mlist[(size_t)(w * B + l) * dc + k]]; }
        std::fill(lane_col.begin(), lane_col.end(), -1); std::fill(bank_col.begin(), bank_col.end(), -1);
        for (int e = 0; e < nE; e++) {
            const int u = eu[e], v = ev[e];
            int a = 0; while (lane_col[(size_t)u * dc + a] >= 0) a++;       // colour free at the lane
            int b = 0; while (bank_col[(size_t)v * dc + b] >= 0) b++;       // colour free at the bank
            if (a != b) {
                // flip the a/b alternating path that starts at bank v with colour a
                std::vector<int> path;
                int node = v; bool at_bank = true; int c = a;
                for (;;) {
                    const int f = at_bank ? bank_col[(size_t)node * dc + c] : lane_col[(size_t)node * dc + c];
                    if (f < 0) break;
                    path.push_back(f);
                    node = at_bank ? eu[f] : ev[f]; at_bank = !at_bank; c = (c == a) ? b : a;
                }
                for (int f : path) { lane_col[(size_t)eu[f] * dc + ecol[f]] = -1; bank_col[(size_t)ev[f] * dc + ecol[f]] = -1; }
                for (int f : path) { ecol[f] = (ecol[f] == a) ? b : a; }
                for (int f : path) { lane_col[(size_t)eu[f] * dc + ecol[f]] = f; bank_col[(size_t)ev[f] * dc + ecol[f]] = f; }
            }
            ecol[e] = a; lane_col[(size_t)u * dc + a] = e; bank_col[(size_t)v * dc + a] = e;
        }
        for (int e = 0; e < nE; e++) {
            if (ecol[e] < 0 || ecol[e] >= dc) return rs;
            rs.order[(size_t)(w * B + eu[e]) * dc + ecol[e]] = e % dc;
        }
        // verify: every step is a permutation of banks
        for (int t = 0; t < dc; t++) {
            unsigned seen = 0;
            for (int l = 0; l < rows; l++) {
                const int k = rs.order[(size_t)(w * B + l) * dc + t];
                if (k < 0) return rs;
                const unsigned bit = 1u << bank[mlist[(size_t)(w * B + l) * dc + k]];
                if (seen & bit) return rs;
                seen |= bit;
            }
            if (rows == B && seen != 0xffffffffu) return rs;
        }
    }
    rs.ok = true;
    return rs;
}

// ---------------------------------------------------------------------------------------------------------------------------
// The same idea for 16-byte accesses and irregular rows (ldpc_ms_quad.cuh).  A 16-byte shared-memory access is served in four
// phases of eight consecutive lanes, and a phase is conflict-free when its eight lanes fall into eight distinct 16-byte bank
// groups.  Thread t owns row row_of_thread[t] (rows sorted by weight, heaviest first, so that the eight rows of a group -- and the
// 32 of a warp -- have the same weight wherever the weight distribution allows); group g runs D_g = its heaviest row's weight
// steps; a lighter row idles (a padding edge) in the steps the colouring leaves it without an edge.
struct GroupSchedule {
    bool ok = false;
    std::vector<int> col, var_of_col;   // [N]
    std::vector<int> row_of_thread;     // [M]
    std::vector<int> step;              // [M*dcm] for thread t, step s: the slot k of its row visited at that step, or -1 (idle)
    int conflicts = 0;                  // edges scheduled with a (two-way) conflict because the bank-group balance was not reached
};

inline GroupSchedule build_group_schedule(int N, int M, int dcm, const std::vector<int> &row_deg, const std::vector<int> &mlist /* [M*dcm] */,
                                          const int B = 8, unsigned seed = 12345)
{
    GroupSchedule gs;
    if (N % B || M < 1 || dcm < 1) return gs;
    gs.row_of_thread.resize(M);
    std::iota(gs.row_of_thread.begin(), gs.row_of_thread.end(), 0);
    std::stable_sort(gs.row_of_thread.begin(), gs.row_of_thread.end(), [&](int a, int b) { return row_deg[a] > row_deg[b]; });
    const int G = (M + B - 1) / B;
    std::vector<int> cap(G, 0);
    for (int t = 0; t < M; t++) cap[t / B] = std::max(cap[t / B], row_deg[gs.row_of_thread[t]]);
    // ---- 1. bank group of every variable: no group of rows may send more than cap[g] of its edges to one bank group ----
    std::vector<std::vector<int>> groups_of(N);
    for (int t = 0; t < M; t++) { const int j = gs.row_of_thread[t]; for (int k = 0; k < row_deg[j]; k++) groups_of[mlist[(size_t)j * dcm + k]].push_back(t / B); }
    for (int i = 0; i < N; i++) if (groups_of[i].size() > 15) return gs;
    std::vector<int> bank(N);
    for (int i = 0; i < N; i++) bank[i] = i % B;
    std::vector<int> load((size_t)G * B, 0);
    for (int i = 0; i < N; i++) for (int g : groups_of[i]) load[(size_t)g * B + bank[i]]++;
    auto sq = [&](int g, int v) { const long long d = v > cap[g] ? v - cap[g] : 0; return d * d; };
    long long cost = 0;
    for (int g = 0; g < G; g++) for (int b = 0; b < B; b++) cost += sq(g, load[(size_t)g * B + b]);
    std::mt19937 rng(seed);
    auto delta_swap = [&](int a, int b) {
        const int ba = bank[a], bb = bank[b];
        long long d = 0;
        int tw[64], tb[64], tv[64], n = 0;
        auto add = [&](int w, int bk, int v) {
            for (int q = 0; q < n; q++) if (tw[q] == w && tb[q] == bk) { tv[q] += v; return; }
            tw[n] = w; tb[n] = bk; tv[n] = v; n++;
        };
        for (int w : groups_of[a]) { add(w, ba, -1); add(w, bb, +1); }
        for (int w : groups_of[b]) { add(w, bb, -1); add(w, ba, +1); }
        for (int q = 0; q < n; q++) { const int cur = load[(size_t)tw[q] * B + tb[q]]; d += sq(tw[q], cur + tv[q]) - sq(tw[q], cur); }
        return d;
    };
    auto do_swap = [&](int a, int b) {
        const int ba = bank[a], bb = bank[b];
        for (int w : groups_of[a]) { load[(size_t)w * B + ba]--; load[(size_t)w * B + bb]++; }
        for (int w : groups_of[b]) { load[(size_t)w * B + bb]--; load[(size_t)w * B + ba]++; }
        bank[a] = bb; bank[b] = ba;
    };
    // focused descent: one of the two variables is taken from an overloaded (group, bank group) pair.  A few overloads may remain
    // (the groups' constraints are tight: cap[g] edges in each of the B bank groups); their edges are scheduled with a conflict.
    std::vector<int> members;
    for (long long it = 0; it < 1500LL * 1000 && cost > 0; it++) {
        int a;
        if ((it & 3) != 3) {
            const int g = (int)(rng() % G), b0 = (int)(rng() % B);
            int vb = -1;
            for (int q = 0; q < B && vb < 0; q++) { const int b = (b0 + q) % B; if (load[(size_t)g * B + b] > cap[g]) vb = b; }
            if (vb < 0) continue;
            members.clear();
            for (int l = 0; l < B && g * B + l < M; l++) {
                const int j = gs.row_of_thread[g * B + l];
                for (int k = 0; k < row_deg[j]; k++) if (bank[mlist[(size_t)j * dcm + k]] == vb) members.push_back(mlist[(size_t)j * dcm + k]);
            }
            a = members[rng() % members.size()];
        } else a = (int)(rng() % N);
        const int b = (int)(rng() % N);
        if (bank[a] == bank[b]) continue;
        const long long d = delta_swap(a, b);
        if (d < 0 || (d == 0 && (rng() & 3) == 0)) { do_swap(a, b); cost += d; }
    }
    gs.col.assign(N, -1); gs.var_of_col.assign(N, -1);
    std::vector<int> fill(B, 0);
    for (int i = 0; i < N; i++) { const int c = bank[i] + B * fill[bank[i]]++; gs.col[i] = c; gs.var_of_col[c] = i; }
    // ---- 2. per group: lane x bank bipartite multigraph of maximum degree <= cap[g] -> cap[g] matchings (Koenig) ----
    gs.step.assign((size_t)M * dcm, -1);
    int clashes = 0;
    for (int g = 0; g < G; g++) {
        const int rows = std::min(B, M - g * B), D = cap[g];
        std::vector<int> eu, ev, ek;                                                   // edge: lane, bank, slot
        for (int l = 0; l < rows; l++) { const int j = gs.row_of_thread[g * B + l]; for (int k = 0; k < row_deg[j]; k++) { eu.push_back(l); ev.push_back(bank[mlist[(size_t)j * dcm + k]]); ek.push_back(k); } }
        const int nE = (int)eu.size();
        {   // an edge beyond cap[g] in its bank group is coloured as if it sat in a bank group that has room (it will conflict there)
            std::vector<int> cnt(B, 0);
            for (int e = 0; e < nE; e++) cnt[ev[e]]++;
            for (int e = 0; e < nE; e++) if (cnt[ev[e]] > D) {
                int to = 0; while (to < B && cnt[to] >= D) to++;
                if (to >= B) return gs;
                cnt[ev[e]]--; cnt[to]++; ev[e] = to; gs.conflicts++;
            }
        }
        std::vector<int> ecol(nE, -1), lane_col((size_t)B * D, -1), bank_col((size_t)B * D, -1);
        for (int e = 0; e < nE; e++) {
            const int u = eu[e], v = ev[e];
            int a = 0; while (a < D && lane_col[(size_t)u * D + a] >= 0) a++;
            int b = 0; while (b < D && bank_col[(size_t)v * D + b] >= 0) b++;
            if (a >= D || b >= D) return gs;
            if (a != b) {                                                              // flip the a/b alternating path that starts at bank v with colour a
                std::vector<int> path;
                int node = v; bool at_bank = true; int c = a;
                for (;;) {
                    const int f = at_bank ? bank_col[(size_t)node * D + c] : lane_col[(size_t)node * D + c];
                    if (f < 0) break;
                    path.push_back(f);
                    node = at_bank ? eu[f] : ev[f]; at_bank = !at_bank; c = (c == a) ? b : a;
                }
                for (int f : path) { lane_col[(size_t)eu[f] * D + ecol[f]] = -1; bank_col[(size_t)ev[f] * D + ecol[f]] = -1; }
                for (int f : path) ecol[f] = (ecol[f] == a) ? b : a;
                for (int f : path) { lane_col[(size_t)eu[f] * D + ecol[f]] = f; bank_col[(size_t)ev[f] * D + ecol[f]] = f; }
            }
            ecol[e] = a; lane_col[(size_t)u * D + a] = e; bank_col[(size_t)v * D + a] = e;
        }
        for (int e = 0; e < nE; e++) gs.step[(size_t)(g * B + eu[e]) * dcm + ecol[e]] = ek[e];
        for (int s = 0; s < D; s++) {                                                  // verify: the lanes of a step touch distinct bank groups
            unsigned seen = 0;
            for (int l = 0; l < rows; l++) {
                const int k = gs.step[(size_t)(g * B + l) * dcm + s];
                if (k < 0) continue;
                const unsigned bit = 1u << bank[mlist[(size_t)gs.row_of_thread[g * B + l] * dcm + k]];
                if (seen & bit) clashes++;
                seen |= bit;
            }
        }
        for (int l = 0; l < rows; l++) {                                               // verify: every edge of the row is visited exactly once
            const int j = gs.row_of_thread[g * B + l];
            unsigned seen = 0;
            for (int s = 0; s < dcm; s++) { const int k = gs.step[(size_t)(g * B + l) * dcm + s]; if (k >= 0) { if (k >= row_deg[j] || (seen >> k) & 1u) return gs; seen |= 1u << k; } }
            if (seen != ((row_deg[j] >= 32) ? 0xffffffffu : ((1u << row_deg[j]) - 1u))) return gs;
        }
    }
    if (clashes > gs.conflicts) return gs;
    gs.ok = true;
    return gs;
}

} // namespace ldpc
