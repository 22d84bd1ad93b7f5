// fnn_relaxed_sm.h — the control flow of NeighborNetLocal.findNodes (NeighborNetLocal.java:170-264) as a resumable
// state machine, so that the whole Relaxed selection can run inside ONE device kernel: a single lane advances the
// machine, and whenever it needs a row scan (findRowMin, :88-157) or an -additive look-ahead (:223-255) it yields to
// the thread block, which does the data-parallel part and resumes it.
//
// The header is plain C++ with host/device qualifiers: csrc/fnn_modes.cuh drives it on the GPU; the CPU tests compile
// the very same code with g++ against the literal oracle (oracle/relaxed_sm_check.cpp, tests/test_relaxed_sm_cpu.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FNN_HD __host__ __device__
#else
#define FNN_HD
#endif

namespace relaxed {

enum Request { REQ_DONE = 0, REQ_SCAN = 1, REQ_LOOKAHEAD = 2, REQ_ERROR = 3 };

// java.util.Random.nextInt(bound) on a 48-bit state
FNN_HD inline int jr_next(unsigned long long& s, int bits) {
    s = (s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int)((long long)s >> (48 - bits));
}
FNN_HD inline int jr_next_int(unsigned long long& s, int bound) {
    int r = jr_next(s, 31);
    const int m = bound - 1;
    if ((bound & m) == 0) return (int)(((long long)bound * (long long)r) >> 31);
    for (int u = r;; u = jr_next(s, 31)) {
        r = u % bound;
        if ((int)((unsigned)u - (unsigned)r + (unsigned)m) >= 0) break;
    }
    return r;
}

// Persistent across iterations (the Java object's fields) plus per-call scratch.  All arrays live in global memory.
struct Machine {
    // ---- persistent (NeighborNetLocal.java:16-19)
    unsigned long long rng;
    int top;
    int first_time;
    int* rowPerm;        // [ntax]
    // ---- per-call storage
    int* cache_epoch;    // [ntax] by position: epoch stamp of the cached list (HashMap foundRowMinimums, identity keys)
    int* cache_list;     // [ntax] by position: list index
    int* list_off;       // [max_lists] offset into tiepool
    int* list_cnt;       // [max_lists]
    int* list_me;        // [max_lists] position of the scanned node (RowMinimum.me)
    int* tiepool;        // [tie_cap] positions of the tied minimisers, position order
    int* mymin;          // [2 * mymin_cap] (me, row) of myMinimums
    int max_lists, tie_cap, mymin_cap;
    int epoch;           // bumped per findNodes call
    int n_lists, tie_used, n_mymin;
    // ---- resumable loop state
    int state;           // see S_* below
    int i;               // outer loop variable (:185)
    int p;               // position of the sampled representative
    int l1, a;           // list of p and the index of the entry being matched (:205)
    int req_pos;         // REQ_SCAN: position to scan
    int pending_key;     // position the requested list will be cached under
    int cx_pos, cy_pos;  // result / last tried candidate (Cx = combineMe.me, Cy = combineMe.row)
    int look_test_pos;   // REQ_LOOKAHEAD: test node position
    int additive;
    int error;
};

enum { S_START = 0, S_OUTER = 1, S_WAIT_L1 = 2, S_MATCH = 3, S_WAIT_L2 = 4, S_DECIDE = 5, S_WAIT_LOOK = 6, S_FINISHED = 7 };

// NodeView: int m(); int id_at(int pos); int nbr_pos(int pos) (-1: singleton)
template <class NodeView>
FNN_HD inline int cached_list(const Machine& M, const NodeView& nv, int pos) {
    if (M.cache_epoch[pos] == M.epoch) return M.cache_list[pos];                 // found.containsKey(p)
    const int nb = nv.nbr_pos(pos);
    if (nb >= 0 && M.cache_epoch[nb] == M.epoch) return M.cache_list[nb];       // found.containsKey(p.nbr)
    return -1;
}

// The caller has written `count` tie positions (position order) at tiepool + tie_used; commit them as the list of
// pending_key.  Returns false on pool overflow.
FNN_HD inline bool commit_scan(Machine& M, int count) {
    if (M.n_lists >= M.max_lists || M.tie_used + count > M.tie_cap) { M.error = 1; return false; }
    const int li = M.n_lists++;
    M.list_off[li] = M.tie_used;
    M.list_cnt[li] = count;
    M.list_me[li] = M.pending_key;
    M.tie_used += count;
    M.cache_epoch[M.pending_key] = M.epoch;
    M.cache_list[M.pending_key] = li;
    return true;
}
FNN_HD inline int tie_room(const Machine& M) { return M.tie_cap - M.tie_used; }

FNN_HD inline void begin_call(Machine& M, int ntax) {
    M.epoch += 1;
    M.n_lists = 0; M.tie_used = 0; M.n_mymin = 0;
    if (M.first_time) {   // :176-182
        for (int k = 0; k < ntax; ++k) M.rowPerm[k] = k;
        M.first_time = 0;
        M.top = ntax - 1;
    }
    M.i = M.top + 1;
    M.state = S_OUTER;
    M.error = 0;
}

// Advance until the machine needs something from the block.  `look_accept`: result of the last look-ahead (S_WAIT_LOOK).
template <class NodeView>
FNN_HD inline Request step(Machine& M, const NodeView& nv, int look_accept) {
    const int m = nv.m();
    while (true) {
        switch (M.state) {
        case S_OUTER: {
            // for (int i = top+1; i > 0; i--)   (:185)  -- `continue` runs the i-- below
            if (!(M.i > 0)) { M.state = S_FINISHED; return REQ_DONE; }
            const int i = M.i;
            const int swapCell = jr_next_int(M.rng, i);
            if (M.rowPerm[swapCell] >= m) {   // stale entry (:187-197)
                const int t = M.rowPerm[swapCell]; M.rowPerm[swapCell] = M.rowPerm[M.top]; M.rowPerm[M.top] = t;
                if (i == M.top + 1) M.i = i - 1; else M.i = i + 1;
                M.top -= 1;
                M.i -= 1;
                break;
            }
            { const int t = M.rowPerm[i - 1]; M.rowPerm[i - 1] = M.rowPerm[swapCell]; M.rowPerm[swapCell] = t; }
            const int p = M.rowPerm[i - 1];
            const int pn = nv.nbr_pos(p);
            if (pn >= 0 && nv.id_at(pn) < nv.id_at(p)) { M.i -= 1; break; }   // one node per cluster (:201-203)
            M.p = p;
            const int li = cached_list(M, nv, p);
            if (li >= 0) { M.l1 = li; M.a = 0; M.state = S_MATCH; break; }
            M.req_pos = p; M.pending_key = p; M.state = S_WAIT_L1;
            return REQ_SCAN;
        }
        case S_WAIT_L1:   // the block scanned row p and committed the list
            M.l1 = M.n_lists - 1; M.a = 0; M.state = S_MATCH;
            break;
        case S_MATCH: {   // for (RowMinimum myRM : testRowMin)  (:205)
            if (M.a >= M.list_cnt[M.l1]) { M.state = S_DECIDE; break; }
            const int row = M.tiepool[M.list_off[M.l1] + M.a];
            int lo = cached_list(M, nv, row);
            if (lo < 0) { M.req_pos = row; M.pending_key = row; M.state = S_WAIT_L2; return REQ_SCAN; }
            // for (RowMinimum testRM : testOtherRow)  (:207-215)
            const int p = M.p, pn = nv.nbr_pos(p);
            const int me = M.list_me[lo];
            for (int b = 0; b < M.list_cnt[lo]; ++b) {
                const int tr = M.tiepool[M.list_off[lo] + b];
                const int rn = nv.nbr_pos(tr);
                if (tr == p || (rn >= 0 && rn == p) || (rn >= 0 && pn >= 0 && rn == pn) || (pn >= 0 && tr == pn)) {
                    if (M.n_mymin >= M.mymin_cap) { M.error = 2; M.state = S_FINISHED; return REQ_ERROR; }
                    M.mymin[2 * M.n_mymin] = me;
                    M.mymin[2 * M.n_mymin + 1] = tr;
                    M.n_mymin += 1;
                    break;
                }
            }
            M.a += 1;
            break;
        }
        case S_WAIT_L2:   // list of `row` is now cached; re-run the match for the same a
            M.state = S_MATCH;
            break;
        case S_DECIDE: {  // :218-259
            if (M.n_mymin == 0) { M.i -= 1; M.state = S_OUTER; break; }
            const int choice = jr_next_int(M.rng, M.n_mymin);
            M.cx_pos = M.mymin[2 * choice];
            M.cy_pos = M.mymin[2 * choice + 1];
            if (!M.additive) { M.state = S_FINISHED; return REQ_DONE; }   // break outerloop
            // -additive: last active node outside both clusters (intended loop), its representative
            const int cxn = nv.nbr_pos(M.cx_pos), cyn = nv.nbr_pos(M.cy_pos);
            int test = -1;
            for (int j = m - 1; j >= 0; --j) {
                if (j == M.cx_pos || j == M.cy_pos || j == cxn || j == cyn) continue;
                test = j;
                break;
            }
            if (test < 0) { M.state = S_FINISHED; return REQ_DONE; }
            const int tn = nv.nbr_pos(test);
            if (tn >= 0 && nv.id_at(tn) < nv.id_at(test)) test = tn;
            M.look_test_pos = test;
            M.state = S_WAIT_LOOK;
            return REQ_LOOKAHEAD;
        }
        case S_WAIT_LOOK:
            if (look_accept) { M.state = S_FINISHED; return REQ_DONE; }
            M.i -= 1; M.state = S_OUTER;   // every repetition of the reference's choice loop repeats this test: fall through
            break;
        case S_FINISHED:
        default:
            return REQ_DONE;
        }
        if (M.error) return REQ_ERROR;
    }
}

}  // namespace relaxed
