// Host-side check of the product's bit-parallel rules (tacult_b200/csrc/rules.cuh, compiled for the CPU)
// against the scalar oracle (oracle/utac_rules.c): same playout specification, every field compared.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../tacult_b200/csrc/rules.cuh"
extern "C" {
#include "../oracle/utac_rules.h"
}

using namespace tc;

static int check_position(const utac_state &o, const Pos &p, const char *where)
{
    uint32_t w[8], ow[8];
    pack(p, w);
    utac_pack_canonical(&o, ow);
    if (memcmp(w, ow, sizeof w)) { printf("state mismatch at %s\n", where); return 1; }
    Summary s = summarize(p);
    if (s.won_mine != o.main_board || (s.won_mine | s.won_opp) != o.main_occ || s.closed != o.game_occ) {
        printf("summary mismatch at %s\n", where); return 1;
    }
    uint32_t lw[3], am[3];
    legal_words(p, s, lw);
    to_action_order(lw, am);
    uint8_t mask[81];
    utac_valid_mask(&o, mask);
    for (int a = 0; a < 81; ++a) {
        bool mine = (am[a >> 5] >> (a & 31)) & 1u;
        if (mine != (mask[a] != 0) || mine != action_bit(lw, a)) { printf("legal mismatch a=%d at %s\n", a, where); return 1; }
    }
    int st = status_of(s);
    int term = utac_is_terminal(&o), tv = utac_terminal_value(&o);
    int expect = !term ? 0 : (tv == 0 ? 2 : (tv == 1 ? 3 : 1));
    if (st != expect) { printf("status mismatch %d vs %d at %s\n", st, expect, where); return 1; }
    float obs[324];
    utac_obs(&o, obs);
    for (int a = 0; a < 81; ++a) {
        bool m = action_bit(p.mine, a), oc = action_bit(p.occ, a);
        if ((obs[a] != 0) != m || (obs[81 + a] != 0) != (oc && !m) || (obs[162 + a] != 0) != (mask[a] != 0) || obs[243 + a] != 1.f) {
            printf("obs mismatch a=%d at %s\n", a, where); return 1;
        }
    }
    return 0;
}

int main(int argc, char **argv)
{
    int games = argc > 1 ? atoi(argv[1]) : 20000;
    uint64_t seed = argc > 2 ? strtoull(argv[2], nullptr, 10) : 12345;
    long plies = 0;
    for (int g = 0; g < games; ++g) {
        utac_state o;
        utac_init(&o);
        Pos p;
        memset(&p, 0, sizeof p);
        if (check_position(o, p, "start")) return 1;
        int ply = 0;
        while (!utac_is_terminal(&o)) {
            uint8_t mask[81];
            int n = utac_valid_mask(&o, mask);
            uint64_t r = utac_mix64(seed ^ ((uint64_t)g * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)ply * 0xD1B54A32D192ED03ull));
            int k = (int)(r % (uint64_t)n), a = -1;
            for (int i = 0; i < 81; ++i) if (mask[i] && k-- == 0) { a = i; break; }
            uint32_t lw[3], am[3];
            legal_words(p, summarize(p), lw);
            to_action_order(lw, am);
            if (nth_action(am, (int)(r % (uint64_t)n)) != a) { printf("nth_action mismatch\n"); return 1; }
            utac_make_move(&o, a);
            if (o.side != 1) utac_flip(&o);
            p = play(p, a);
            ++ply;
            ++plies;
            if (check_position(o, p, "playout")) { printf("game %d ply %d\n", g, ply); return 1; }
        }
    }
    printf("OK games=%d plies=%ld\n", games, plies);
    return 0;
}
