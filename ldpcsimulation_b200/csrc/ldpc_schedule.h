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

} // namespace ldpc
