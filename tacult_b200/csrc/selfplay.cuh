// Device-resident self-play driver: `Coach.executeEpisode` semantics
// (/root/reference/src/tacult/base/coach.py:96-157) without host round trips.
#pragma once
#include "tree.cuh"

namespace tc {

struct EnvState {
    uint32_t state[8];     // canonical position of the side to move
    int32_t player;        // +1 first player, -1 second (coach.py:90)
    int32_t ply;           // moves played in this game == _episodeStep before increment (coach.py:124)
    int32_t game;          // per-env game ordinal
    int32_t done;          // 1: game over and no auto-reset -> env idles (coach.py:133-134)
    int32_t needs_reset;   // 1: tree must be cleared before the next search
    int32_t n_records;     // records of the current game held in `pending`
    int32_t pad[2];
};

struct SelfPlayStore {
    int E;
    EnvState *env;
    tc_example *pending;           // [E][81] records of the games in progress
    tc_example *finished;          // [finished_cap] records of completed games
    int finished_cap;
    int *finished_count;
    unsigned long long *stats;     // plies, sims, games, first wins, second wins, draws, dropped records
    uint32_t *root_states;         // [E][8] scratch handed to k_set_roots
    uint8_t *root_active;          // [E]
};

void launch_selfplay_init(const SelfPlayStore &sp, cudaStream_t st);
void launch_selfplay_roots(const SelfPlayStore &sp, cudaStream_t st);
void launch_selfplay_move(const TreeStore &ts, const SelfPlayStore &sp, tc_selfplay_config cfg, cudaStream_t st);
void launch_examples_from_records(int n_out, const tc_example *records, int n_records, const int32_t *index, float *obs,
                                  float *pi, float *z, cudaStream_t st);
void launch_reset_marked(const TreeStore &ts, const SelfPlayStore &sp, cudaStream_t st);

}  // namespace tc
