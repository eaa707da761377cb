/*
 * tacult_b200 — C ABI of the B200-native self-play hot path.
 *
 * The reference (lunathanael/tacult) has no FFI of its own: its hot path is three
 * duck-typed Python protocols.  Each entry point below names the
 * reference interface it stands in for; the Python modules under tacult_b200/ bind them with ctypes and
 * re-creates those Python protocols on top (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns TC_OK (0) or a negative TC_E* code; tc_last_error()
 *     returns a human-readable message for the calling thread's last failure;
 *   - pointers named *_host are host buffers owned by the caller, *_dev are device
 *     pointers on the engine's device; no torch types cross this boundary;
 *   - an engine handle is not thread-safe; all work is issued on the engine's CUDA
 *     stream (tc_set_stream) and the host-buffer calls synchronise before returning.
 *
 * Position format ("canonical frame", what VectorizedMCTS and the network always see:
 * side to move == 1 and its stones in `board`, /root/reference/src/tacult/utac_game.py:83-107)
 *   fields form : uint16 board[9], uint16 occ[9], int8 last_square   — GAMESTATE members,
 *                 bit k of sub-board i is global cell (3*(i/3)+k/3, 3*(i%3)+k%3)
 *                 (utac_game.py:259-265,283-291)
 *   packed form : 8 x uint32: w[g] (g=0..2) = board[3g] | board[3g+1]<<9 | board[3g+2]<<18,
 *                 w[3+g] = the same for occ, w[6] bits 0..3 = last_square+1, w[7] = 0.
 */
#ifndef TACULT_B200_H
#define TACULT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TC_ABI_VERSION 2
#define TC_ACTIONS 81
#define TC_STATE_WORDS 8

enum {
    TC_OK = 0,
    TC_EINVAL = -1,     /* bad argument */
    TC_ECUDA = -2,      /* CUDA runtime / driver failure */
    TC_ECAPACITY = -3,  /* a tree ran out of node / edge / hash capacity */
    TC_ESTATE = -4,     /* call sequence error (e.g. search before weights are loaded) */
    TC_EILLEGAL = -5    /* an illegal move was requested */
};

enum {
    TC_EVAL_NET_F32 = 0,   /* fp32 CUDA-core forward: parity mode, 1e-3 of the reference net */
    TC_EVAL_NET_BF16 = 1,  /* bf16 tcgen05/TMEM trunk: production mode, 2e-2 */
    TC_EVAL_HASH = 2,      /* deterministic integer-hash priors/values: bit-exact visit-count tests */
    TC_EVAL_NET_BF16_MASKED = 3   /* bf16 chain whose policy head masks the soft-max with the leaf's legal moves and
                                     renormalises in float32 (the composition mcts.py:272-275 applies in float64 to the
                                     unmasked output); the tree stores those priors as they come.  Production mode:
                                     2e-2 of the reference priors, not bit-identical to parity mode. */
};

/* terminal status of a canonical position, seen by its side to move */
enum { TC_LIVE = 0, TC_LOST = 1, TC_DRAW = 2, TC_WON = 3 };

typedef struct tc_engine tc_engine;

typedef struct {
    int32_t num_envs;            /* trees searched in lock-step == args.numEnvs (mcts.py:152-162) */
    int32_t max_nodes_per_tree;  /* node pool per tree; the reference never prunes inside a game */
    int32_t max_edges_per_tree;  /* legal-move edge pool per tree; 0 = 12 x max_nodes_per_tree */
    int32_t device;              /* CUDA device ordinal */
    int32_t evaluator;           /* TC_EVAL_* */
    int32_t eval_log_capacity;   /* >0: keep the first N leaf evaluations (state, pi, v) for replay tests */
    uint64_t hash_salt;          /* TC_EVAL_HASH parameters (oracle/evaluators.py) */
    int32_t hash_pi_levels;
    int32_t hash_v_levels;
} tc_config;

/* shape of `_UtacNNet` (/root/reference/src/tacult/network.py:44-79) */
typedef struct {
    int32_t channels;
    int32_t num_residual_blocks;
    int32_t embedding_size;
    int32_t in_planes;           /* 4 in the reference (network.py:55) */
} tc_net_desc;

/* counters SURVEY.md §8(d) asks the kernels to export (summed over trees since the last reset) */
typedef struct {
    uint64_t sims;               /* simulations started from a live root */
    uint64_t sum_depth;          /* interior levels traversed (PUCT selections) */
    uint64_t sum_legal;          /* legal edges scanned by those selections */
    uint64_t nn_leaves;          /* leaves sent to the evaluator */
    uint64_t sum_leaf_legal;     /* legal moves of those leaves (edges written at expansion) */
    uint64_t terminal_leaves;    /* simulations that ended in a finished game */
    uint64_t transpositions;     /* first visits of an edge whose successor already existed */
    uint64_t nodes_in_use;       /* node pool entries in use, all trees */
    uint64_t edges_in_use;
    uint64_t max_nodes_one_tree;
    uint64_t max_edges_one_tree;
    uint64_t kernel_launches;    /* kernels this library launched since the last tc_reset_stats */
} tc_stats;

int         tc_abi_version(void);
/* free / total bytes of HBM on `device` (used by the host side to size tree pools) */
int         tc_device_memory(int device, uint64_t *free_bytes, uint64_t *total_bytes);
const char *tc_last_error(void);
/* page-locked host memory for the *_host buffers of this API (optional: any host pointer works, pinned ones copy faster
 * and without a staging pass); free with tc_host_free */
int         tc_host_alloc(size_t bytes, void **out);
int         tc_host_free(void *p);

/* ---- rules engine hooks: replace utac_gym.core.GameState methods (external C++ `utac`;
 *      call sites /root/reference/src/tacult/utac_game.py:45-46,62,75-81,248) -------------- */
/* make_move + getCanonicalForm: out[i] = canonical successor of states[i] after actions[i];
 * status[i] = TC_* of the successor (utac_game.py:34-47,64-107).  Illegal action -> TC_EILLEGAL. */
int tc_rules_step(int n, const uint32_t *states_host, const int8_t *actions_host,
                  uint32_t *out_states_host, int8_t *status_host);
/* get_valid_mask (utac_game.py:62): mask[i][a] in {0,1}, action order */
int tc_rules_legal_mask(int n, const uint32_t *states_host, uint8_t *mask_host);
/* is_terminal/terminal_value (utac_game.py:64-81) */
int tc_rules_status(int n, const uint32_t *states_host, int8_t *status_host);
/* get_obs (utac_game.py:247-250, base/utils.py:7-10): obs[i][4][9][9] float32 */
int tc_rules_encode_obs(int n, const uint32_t *states_host, float *obs_host);
/* seeded random playouts on the device, spec shared with oracle/utac_rules.c:utac_playout */
typedef struct {
    uint64_t checksum;
    uint32_t final_state[TC_STATE_WORDS];
    int32_t plies;
    int32_t outcome;
} tc_playout_result;
int tc_rules_playouts(uint64_t seed, uint64_t first_game, int n, int max_plies, tc_playout_result *out_host);

/* ---- host-side game-state object: utac_gym.core.GameState / utac_gym.core.types.GAMESTATE (external C++ `utac`
 *      behind nanobind in the reference; surface inferred from the call sites).  Plain host
 *      functions on a caller-owned struct, no device involved: the reference keeps these objects on the host too.
 *      board = stones of player +1, occ = all stones, side = player to move (+1 / -1), last_square = cell of the
 *      last move (-1 at game start); main_board / main_occ / game_occ are derived (utac_game.py:222-229). ---------- */
typedef struct {
    uint16_t board[9];           /* GAMESTATE.board  (utac_game.py:101-103) */
    uint16_t occ[9];             /* GAMESTATE.occ */
    uint16_t main_board;         /* sub-boards won by player +1 (utac_game.py:104) */
    uint16_t main_occ;           /* sub-boards won by either player */
    uint16_t game_occ;           /* sub-boards closed: won or full */
    int8_t side;                 /* utac_game.py:105-106 */
    int8_t last_square;          /* utac_game.py:182-184,215-216 */
} tc_gamestate;
void tc_gs_init(tc_gamestate *s);                                  /* GameState()            utac_game.py:18 */
void tc_gs_rederive(tc_gamestate *s);                              /* GameState(GAMESTATE)   utac_game.py:107,218 */
int  tc_gs_make_move(tc_gamestate *s, int action);                 /* .make_move             utac_game.py:46; TC_EILLEGAL */
int  tc_gs_valid_mask(const tc_gamestate *s, double *mask81);      /* .get_valid_mask        utac_game.py:62; returns #legal */
int  tc_gs_is_terminal(const tc_gamestate *s);                     /* .is_terminal           utac_game.py:75 */
int  tc_gs_terminal_value(const tc_gamestate *s);                  /* .terminal_value        utac_game.py:78,81 */
void tc_gs_obs(const tc_gamestate *s, float *obs324);              /* .get_obs               utac_game.py:248, base/utils.py:8 */
/* the canonicalBoards list of getActionProbs (mcts.py:175) -> packed roots in one call; active[i] == 0 on entry marks a
 * None board (mcts.py:235-239); finished boards whose side is not +1 (arena.py:176-184) come back inactive */
int  tc_gs_pack(int n, const tc_gamestate *gs_host, uint32_t *states_host, uint8_t *active_host);
/* The actor's move choice from the root visit counts of tc_root_counts, on the host, one call for all envs
 * (base/coach.py:125,131,140: temps = episodeStep < tempThreshold, pi = getActionProbs(...), action = np.random.choice(81, p=pi);
 * base/mcts.py:200-214: tau = 1 -> pi = counts / sum, tau = 0 -> the arg-max).  counts_host u32[n][81]; tau_one u8[n] (1: sample
 * a ~ counts / sum with the caller's uniform u[i] in [0,1): the first action whose cumulative count exceeds floor(u * sum);
 * 0: arg-max, lowest index on ties); actions_host i8[n].  A row of zeros (finished env) gives action 0. */
int  tc_choose_actions(int n, const uint32_t *counts_host, const uint8_t *tau_one_host, const double *u_host,
                       int8_t *actions_host);

/* ---- search engine: replaces VectorizedMCTS (/root/reference/src/tacult/base/mcts.py:147-358) ---- */
int tc_create(const tc_config *cfg, tc_engine **out);              /* VectorizedMCTS.__init__ mcts.py:152-162 */
int tc_destroy(tc_engine *e);
int tc_set_stream(tc_engine *e, void *cuda_stream);                /* cudaStream_t; NULL = engine-owned stream */
int tc_sync(tc_engine *e);                                         /* waits for everything enqueued on that stream */
int tc_reset(tc_engine *e, int env);                               /* VectorizedMCTS.reset(i) mcts.py:165-172; env<0 = all */
int tc_reset_envs(tc_engine *e, int n, const int32_t *envs_host);  /* the same for a list of envs, one launch */
int tc_set_hash_evaluator(tc_engine *e, uint64_t salt, int pi_levels, int v_levels);
int tc_set_evaluator(tc_engine *e, int evaluator);
/* roots for the next tc_search: canonicalBoards of getActionProbs (mcts.py:175).  active[i]==0 marks a
 * None board (mcts.py:235-239).  Terminal boards are accepted and simply do nothing (mcts.py:250-258). */
int tc_set_roots(tc_engine *e, const uint16_t *board_host, const uint16_t *occ_host,
                 const int8_t *last_square_host, const uint8_t *active_host);
int tc_set_roots_packed(tc_engine *e, const uint32_t *states_host, const uint8_t *active_host);
/* One ply for every env on the device: the env's root (as last set by tc_set_roots* or tc_advance) + actions[i] -> the
 * canonical successor (getNextState + getCanonicalForm, utac_game.py:34-47,83-107), which becomes the root of the next
 * tc_search; the tree and its statistics are kept, as the reference's dicts are across plies (coach.py:162).
 * actions[i] < 0 leaves env i where it is.  restart != 0: an env whose game just ended gets a fresh tree and the initial
 * position (VectorizedMCTS.reset(i) mcts.py:165-172 + getInitBoard), otherwise it idles.  states_host / status_host
 * (optional): the successors and their TC_* status as seen by their side to move, before any restart.
 * An illegal action -> TC_EILLEGAL (that env keeps its position). */
int tc_advance(tc_engine *e, const int8_t *actions_host, int restart, uint32_t *states_host, int8_t *status_host);
/* num_sims x VectorizedMCTS.search (mcts.py:184-185,217-342); cpuct read per call like args.cpuct (mcts.py:309) */
int tc_search(tc_engine *e, int num_sims, double cpuct);
/* root edge visit counts, the `_counts` of getActionProbs (mcts.py:191-198): counts[i][a] */
int tc_root_counts(tc_engine *e, uint32_t *counts_host);
/* root edge mean values Q (mcts.py:156), NaN where the edge was never visited; test hook */
int tc_root_q(tc_engine *e, double *q_host);
int tc_get_stats(tc_engine *e, tc_stats *out);
int tc_reset_stats(tc_engine *e);
/* per-kernel-class device time, measured with CUDA events around every launch on the engine's stream
 * (adds two event records per launch: enable only for profiling passes, not for headline timings) */
enum {
    TC_K_SELECT = 0, TC_K_EVAL_HASH = 1, TC_K_NN_STEM = 2, TC_K_NN_CONV = 3, TC_K_NN_FC1 = 4, TC_K_NN_HEADS = 5,
    TC_K_EXPAND_BACKUP = 6, TC_K_OTHER = 7, TC_K_CLASSES = 8
};
int tc_profile_enable(tc_engine *e, int on);
int tc_profile_read(tc_engine *e, double *ms_by_class, uint64_t *launches_by_class);   /* [TC_K_CLASSES]; resets */
/* leaf evaluations recorded when cfg.eval_log_capacity > 0: returns how many, copies up to `max` */
int tc_eval_log(tc_engine *e, int max, uint32_t *states_host, float *pi_host, float *v_host, int *count_out);

/* ---- evaluator: replaces NNetWrapper.predict + _UtacNNet.forward
 *      (/root/reference/src/tacult/base/nn_wrapper.py:125-136, network.py:81-102) ------------------------ */
/* blob = the reference state_dict tensors, float32, concatenated in this order:
 *   conv_in.{weight,bias}, bn_in.{weight,bias,running_mean,running_var},
 *   for k in blocks: resnet.k.conv_block1.0.{weight,bias}, resnet.k.conv_block1.1.{weight,bias,running_mean,running_var},
 *                    resnet.k.conv_block2.0.{...}, resnet.k.conv_block2.1.{...},
 *   fc1.{weight,bias}, fc_bn1.{weight,bias,running_mean,running_var}, pi.{weight,bias}, v.{weight,bias}
 * BatchNorm (eps 1e-5) is folded into the preceding layer inside the library. */
int tc_load_weights(tc_engine *e, const tc_net_desc *desc, const float *blob_host, size_t n_floats);
size_t tc_weight_blob_floats(const tc_net_desc *desc);
/* predict: obs[B][in_planes][9][9] f32 -> pi[B][81] (unmasked softmax, network.py:99), v[B] (tanh).
 * `evaluator` selects TC_EVAL_NET_F32 or TC_EVAL_NET_BF16.  logits_host may be NULL. */
int tc_forward(tc_engine *e, int evaluator, int batch, const float *obs_host, float *pi_host, float *v_host,
               float *logits_host);
/* the same from packed positions (observation planes built on the device) */
int tc_forward_states(tc_engine *e, int evaluator, int batch, const uint32_t *states_host, float *pi_host,
                      float *v_host);
/* device-resident variant used for throughput sweeps: obs already in HBM as packed states */
int tc_forward_states_dev(tc_engine *e, int evaluator, int batch, const uint32_t *states_dev, float *pi_dev,
                          float *v_dev);

/* the same from dense observation planes already in HBM: obs_dev[B][in_planes][9][9] float32 (any in_planes <= 32, e.g. the
 * 8-frame history variant) */
int tc_forward_obs_dev(tc_engine *e, int evaluator, int batch, const float *obs_dev, float *pi_dev, float *v_dev);

/* test hook: trunk activations after conv_in and the first n_convs residual-block convolutions,
 * float32 [batch][81][channels] (NHWC), for either network evaluator */
int tc_debug_trunk(tc_engine *e, int evaluator, int batch, const uint32_t *states_host, int n_convs, float *out_host);

/* ---- device-resident self-play driver: Coach.executeEpisode semantics
 *      (/root/reference/src/tacult/base/coach.py:87-157) without host round trips ------------------------- */
#define TC_SP_DISCARD_RECORDS 1
typedef struct {
    int32_t num_sims;            /* args.numMCTSSims */
    int32_t temp_threshold;      /* args.tempThreshold: tau=1 while episodeStep < threshold (coach.py:124-125) */
    int32_t auto_reset;          /* 1: a finished env restarts a new game with a fresh tree (the reference's
                                    commented-out block coach.py:114-122); 0: it idles like the reference */
    int32_t flags;               /* TC_SP_DISCARD_RECORDS: play, but keep no trajectory records (warm-up, benchmarks) */
    double cpuct;
    uint64_t seed;
} tc_selfplay_config;
typedef struct {
    uint64_t plies;              /* env-plies played (live envs only) */
    uint64_t sims;               /* live envs x num_sims summed over plies == the BASELINE.json metric numerator */
    uint64_t games_finished;
    uint64_t first_player_wins, second_player_wins, draws;
    uint64_t examples;           /* trajectory records available */
} tc_selfplay_stats;
/* one record per env-ply: what coach.py:136-138 stores before symmetry augmentation */
typedef struct {
    uint32_t state[TC_STATE_WORDS];   /* canonical position */
    uint16_t counts[TC_ACTIONS];      /* root visit counts (pi = counts/sum or one-hot, per tau) */
    int8_t action;                    /* move played */
    int8_t tau_one;                   /* 1 if sampled with tau=1 */
    int8_t player;                    /* +1 first player, -1 second */
    int8_t z_sign;                    /* filled at game end: sign of z for `player`, coach.py:143-149 */
    int32_t env;
    int32_t game;                     /* per-env game ordinal */
    int32_t ply;
    int32_t finished;                 /* 0 game in progress; 1 decisive: z = z_sign; 2 draw: z = z_sign * 1e-4 */
} tc_example;
int tc_selfplay_begin(tc_engine *e, const tc_selfplay_config *cfg);
int tc_selfplay_configure(tc_engine *e, const tc_selfplay_config *cfg);   /* change sims/cpuct/... mid-run, no reset */
int tc_selfplay_step(tc_engine *e, int plies);      /* plays `plies` lock-step plies; asynchronous on the stream */
int tc_selfplay_sync(tc_engine *e);
int tc_selfplay_get_stats(tc_engine *e, tc_selfplay_stats *out);
/* copies finished-game records to the host (or leaves them on the device for NCCL): returns count */
int tc_selfplay_drain(tc_engine *e, int max, tc_example *out_host, int *count_out);
int tc_selfplay_drain_dev(tc_engine *e, int max, tc_example *out_dev, int *count_out);
/* Trajectory records -> training examples on the device: what coach.py:136-149 builds on the host — the eight dihedral
 * images of (obs, pi) in getSymmetries' order (utac_game.py:109-220: for i in 0..3, for flip in (False, True)), pi from
 * the counts (tau = 1) or the one-hot of the move played (tau = 0, mcts.py:206-213), z = z_sign (x 1e-4 for a draw).
 * Example j of a record array of n_records is image j % 8 of record j / 8.  Output row r holds example index_dev[r]
 * (index_dev == NULL: r itself): obs[n_out][4][9][9], pi[n_out][81], z[n_out], all float32 device buffers. */
int tc_examples_from_records_dev(tc_engine *e, int n_out, const tc_example *records_dev, int n_records,
                                 const int32_t *index_dev, float *obs_dev, float *pi_dev, float *z_dev);

#ifdef __cplusplus
}
#endif
#endif /* TACULT_B200_H */
