// Device-resident tree store and the warp-per-tree search kernels.
//
// Replaces the six per-env Python dicts of VectorizedMCTS
// (/root/reference/src/tacult/base/mcts.py:152-162) and the per-level Python loops of
// `_search` (mcts.py:243-342).  Semantics kept bit for bit:
//   * a node exists once it has been expanded (`s in Ps`); nodes are found by full state key
//     (board, occ, last_square — utac_game.py:222-241), so transpositions merge exactly as
//     the reference's dict lookups do, and statistics persist across plies;
//   * per edge: prior P (float64), mean value Q (float64, NOT a sum), visit count N;
//     per node: Ns;
//   * PUCT, expansion and backup use float64 with separately rounded operations
//     (__dmul_rn/__dadd_rn/__ddiv_rn/__dsqrt_rn: no FMA contraction).
//
// Layout in HBM (one warp owns one tree, so nothing below needs atomics):
//   node header  32 B : mine[3] occ[3] meta(last_square+1 | n_edges<<8) edge_off   — immutable; only read when
//                       a node is expanded or looked up, never on the way down
//   edge block  32n B : n records of 32 bytes, edges in action order, each two 16-byte halves a lane moves with ONE
//                       vector load / store:  { Q f64, N u32, child u32 } — what the backup reads and rewrites — and
//                       { P f64, coff u32, meta u32 } — written once (creation, expansion).  meta = action |
//                       n_edges(child)<<8, coff = edge block of the child: an edge carries everything needed to step to
//                       the child's edge block, so one level of the descent is ONE dependent memory round trip (two
//                       16-byte loads per lane).  Ns[s] is not stored: it always equals the sum of the node's edge
//                       counts (both are incremented together, mcts.py:333-340).
//   hash table   4 B/slot, per tree, open addressing, entry = tag(12) | node index + 1 (20)
#pragma once
#include "common.cuh"
#include "rules.cuh"

namespace tc {

constexpr uint32_t CHILD_UNRESOLVED = 0xFFFFFFFFu;
constexpr uint32_t CHILD_TERM_BASE = 0xFFFFFFF0u;     // CHILD_TERM_BASE + status (1 lost, 2 draw, 3 won)
constexpr int MAX_DEPTH = 82;
constexpr uint32_t MAX_NODES_PER_TREE = (1u << 20) - 2;
constexpr unsigned FULL = 0xFFFFFFFFu;

enum { LEAF_NONE = 0, LEAF_NN = 1, LEAF_TERMINAL = 2 };

struct TreeCtl {
    uint32_t root_state[8];
    int32_t root_idx;        // -1: root position not (yet) in the tree
    int32_t active;          // env passed a board, the position is not terminal and the pools are not exhausted
    uint32_t n_nodes;
    uint32_t n_edges;
    int32_t leaf_kind;
    int32_t leaf_node;
    int32_t leaf_row;
    int32_t path_len;
    double leaf_value;       // value of a terminal leaf seen by its side to move (mcts.py:254-258)
    uint32_t root_off;       // edge block of the root
    uint32_t root_n;
    // the leaf awaiting evaluation, as the descent left it: its edge block and legal moves (action order), so that the
    // expansion needs neither the node header nor the rules again
    uint32_t leaf_off, leaf_n;
    uint32_t leaf_am[3];
    uint32_t pad_;
    // counters
    unsigned long long sims, sum_depth, sum_legal, nn_leaves, sum_leaf_legal, terminal_leaves, transpositions;
};

struct TreeStore {
    int E;
    uint32_t capN, capE, hashSize;
    uint4 *node_hdr;
    unsigned long long *edge_mem;   // [E][capE][4]: 32 bytes per edge
    uint32_t *hash;
    TreeCtl *ctl;
    uint2 *path;                    // [E][MAX_DEPTH]: (edge_off, slot | n<<8)
    uint32_t *eval_state;    // [E][8]  leaves to evaluate, compacted
    float *eval_pi;          // [E][81]
    float *eval_v;           // [E]
    int *eval_count;
    int *eval_count_next;       // counter of the NEXT simulation, zeroed by this one's descent kernel (or null)
    int *error_flag;
    int pdl_trigger;            // griddepcontrol.launch_dependents at kernel entry (common.cuh)
    int priors_masked;          // eval_pi is already masked + renormalised (TC_EVAL_NET_BF16_MASKED): store it as is
};

struct __align__(16) Edge {        // 32 bytes
    double Q;                      // mean value of the edge (mcts.py:156), meaningful when N > 0
    uint32_t N;                    // visit count
    uint32_t child;                // CHILD_UNRESOLVED, a node index, or CHILD_TERM_BASE + status
    double P;                      // prior (float64, as the reference stores it)
    uint32_t coff;                 // edge block of the child
    uint32_t meta;                 // action | n_edges(child) << 8
};
static_assert(sizeof(Edge) == 32, "edge record is two 16-byte halves");

#ifdef __CUDACC__
// first edge of a tree's pool; a node's block is tree_edges(...) + edge_off
__device__ __forceinline__ Edge *tree_edges(const TreeStore &ts, int tree)
{
    return reinterpret_cast<Edge *>(ts.edge_mem) + (size_t)tree * ts.capE;
}
// the halves as 16-byte vectors: x,y = the double, z,w = the two words
__device__ __forceinline__ uint4 edge_load_stats(const Edge *e) { return *reinterpret_cast<const uint4 *>(e); }
__device__ __forceinline__ uint4 edge_load_link(const Edge *e) { return *(reinterpret_cast<const uint4 *>(e) + 1); }
__device__ __forceinline__ double u2d(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }
__device__ __forceinline__ uint4 edge_pack(double d, uint32_t a, uint32_t b)
{
    return make_uint4((uint32_t)__double2loint(d), (uint32_t)__double2hiint(d), a, b);
}
#endif

struct HashEvalParams {
    unsigned long long salt;
    int pi_levels, v_levels;
};

void launch_set_roots(const TreeStore &ts, const uint32_t *states_dev, const uint8_t *active_dev, cudaStream_t st);
void launch_reset(const TreeStore &ts, int env, cudaStream_t st);
void launch_reset_list(const TreeStore &ts, int n, const int32_t *envs_dev, cudaStream_t st);
void launch_advance(const TreeStore &ts, const int8_t *actions_dev, int restart, uint32_t *states_dev, int8_t *status_dev,
                    int *illegal_dev, cudaStream_t st);
void launch_select(const TreeStore &ts, double cpuct, cudaStream_t st);
void launch_expand_backup(const TreeStore &ts, cudaStream_t st);
void launch_backup_select(const TreeStore &ts, double cpuct, cudaStream_t st);
void launch_hash_eval(const TreeStore &ts, HashEvalParams hp, cudaStream_t st);
void launch_root_counts(const TreeStore &ts, uint32_t *counts_dev, double *q_dev, cudaStream_t st);
void launch_pack_fields(int n, const uint16_t *board, const uint16_t *occ, const int8_t *last, uint32_t *out,
                        cudaStream_t st);

}  // namespace tc
