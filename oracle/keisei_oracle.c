/*
 * keisei_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference engine's algorithm for the self-play
 * rollout hot path (tachyon-beep/shogidrl, "Keisei").  It follows the
 * reference literally -- pseudo-legal generation, simulate-every-candidate and
 * test "is my king attacked" by re-scanning all 81 squares, full nested
 * regeneration for uchifuzume, full-state (not hashed) repetition compare --
 * so that it is algorithmically independent of the CUDA path it checks (which
 * uses checker / pin / danger sets and Zobrist keys).
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 * golden traces produced by importing the Python reference itself
 * (oracle/gen_golden.py, fixtures under tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * All file:line citations are relative to the reference checkout.
 *
 * Piece code used everywhere in this repo: 0 = empty, else
 * 1 + type + 14*color with type per keisei/shogi/shogi_core_definitions.py:64-83
 * (P0 L1 N2 S3 G4 B5 R6 K7 +P8 +L9 +N10 +S11 +B12 +R13) and colour per :50-61
 * (BLACK=0 moves toward row 0, WHITE=1).  Square index = row*9 + col.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define NSQ 81
#define NACT 13527
#define T_PAWN 0
#define T_LANCE 1
#define T_KNIGHT 2
#define T_SILVER 3
#define T_GOLD 4
#define T_BISHOP 5
#define T_ROOK 6
#define T_KING 7
#define T_PPAWN 8
#define T_PLANCE 9
#define T_PKNIGHT 10
#define T_PSILVER 11
#define T_PBISHOP 12
#define T_PROOK 13

#define R_NONE 0
#define R_TSUMI 1      /* "Tsumi"              shogi_core_definitions.py:138 */
#define R_STALEMATE 2  /* "stalemate"          :139 */
#define R_MAXMOVES 3   /* "Max moves reached"  :141 */
#define R_SENNICHITE 4 /* "Sennichite"         :140 */

#define SNAP 98 /* 81 board + 16 hands + 1 side */

typedef struct {
  int8_t board[NSQ];
  int32_t hands[2][8]; /* [colour][type]; slot 7 only via the king-capture quirk */
  int32_t side;
  int32_t move_count;
  int32_t max_moves;
  int32_t game_over;
  int32_t winner; /* -1 = None */
  int32_t reason;
  int32_t hist_len, hist_cap;
  uint8_t *hist; /* move_history state_hash entries, SNAP bytes each */
} orc_game;

static inline int mk(int type, int color) { return 1 + type + 14 * color; }
static inline int ptype(int code) { return (code - 1) % 14; }
static inline int pcolor(int code) { return (code - 1) / 14; }
static inline int on_board(int r, int c) { return r >= 0 && r < 9 && c >= 0 && c < 9; }

/* shogi_core_definitions.py:211-224 */
static inline int promoted_of(int t) {
  switch (t) {
    case T_PAWN: return T_PPAWN;
    case T_LANCE: return T_PLANCE;
    case T_KNIGHT: return T_PKNIGHT;
    case T_SILVER: return T_PSILVER;
    case T_BISHOP: return T_PBISHOP;
    case T_ROOK: return T_PROOK;
  }
  return -1;
}
static inline int base_of(int t) {
  switch (t) {
    case T_PPAWN: return T_PAWN;
    case T_PLANCE: return T_LANCE;
    case T_PKNIGHT: return T_KNIGHT;
    case T_PSILVER: return T_SILVER;
    case T_PBISHOP: return T_BISHOP;
    case T_PROOK: return T_ROOK;
  }
  return t; /* PROMOTED_TO_BASE_TYPE.get(t, t): shogi_move_execution.py:109-111 */
}

/* ---- generate_piece_potential_moves: shogi_rules_logic.py:82-208 ---- */
static int potential_moves(const int8_t *b, int code, int r0, int c0, int *out) {
  int n = 0;
  uint8_t seen[NSQ];
  memset(seen, 0, sizeof seen);
  int color = pcolor(code), t = ptype(code);
  int f = color == 0 ? -1 : 1; /* :96-98 */
  int offs[8][2];
  int no = 0;
#define ADD(dr, dc) do { offs[no][0] = (dr); offs[no][1] = (dc); no++; } while (0)
  if (t == T_PAWN) { ADD(f, 0); }
  else if (t == T_KNIGHT) { ADD(2 * f, -1); ADD(2 * f, 1); }
  else if (t == T_SILVER) { ADD(f, 0); ADD(f, -1); ADD(f, 1); ADD(-f, -1); ADD(-f, 1); }
  else if (t == T_GOLD || t == T_PPAWN || t == T_PLANCE || t == T_PKNIGHT || t == T_PSILVER) {
    ADD(f, 0); ADD(f, -1); ADD(f, 1); ADD(0, -1); ADD(0, 1); ADD(-f, 0);
  } else if (t == T_KING) {
    ADD(-1, -1); ADD(-1, 0); ADD(-1, 1); ADD(0, -1); ADD(0, 1); ADD(1, -1); ADD(1, 0); ADD(1, 1);
  }
  if (t == T_PBISHOP) { ADD(-1, 0); ADD(1, 0); ADD(0, -1); ADD(0, 1); }   /* :176-182 */
  if (t == T_PROOK) { ADD(-1, -1); ADD(-1, 1); ADD(1, -1); ADD(1, 1); }   /* :186-192 */
#undef ADD
  for (int i = 0; i < no; i++) { /* :162-167 */
    int r = r0 + offs[i][0], c = c0 + offs[i][1];
    if (!on_board(r, c)) continue;
    int tc = b[r * 9 + c];
    if (tc == 0 || pcolor(tc) != color) {
      if (!seen[r * 9 + c]) { seen[r * 9 + c] = 1; out[n++] = r * 9 + c; }
    }
  }
  int dirs[4][2];
  int nd = 0;
  if (t == T_LANCE) { dirs[0][0] = f; dirs[0][1] = 0; nd = 1; }
  else if (t == T_BISHOP || t == T_PBISHOP) {
    int d[4][2] = {{-1, -1}, {-1, 1}, {1, -1}, {1, 1}};
    memcpy(dirs, d, sizeof d); nd = 4;
  } else if (t == T_ROOK || t == T_PROOK) {
    int d[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
    memcpy(dirs, d, sizeof d); nd = 4;
  }
  for (int k = 0; k < nd; k++) { /* :194-206 */
    for (int i = 1; i < 9; i++) {
      int r = r0 + dirs[k][0] * i, c = c0 + dirs[k][1] * i;
      if (!on_board(r, c)) break;
      int tc = b[r * 9 + c];
      if (tc == 0) {
        if (!seen[r * 9 + c]) { seen[r * 9 + c] = 1; out[n++] = r * 9 + c; }
      } else {
        if (pcolor(tc) != color && !seen[r * 9 + c]) { seen[r * 9 + c] = 1; out[n++] = r * 9 + c; }
        break;
      }
    }
  }
  return n;
}

/* find_king: shogi_rules_logic.py:25-32 */
static int find_king(const int8_t *b, int color) {
  for (int s = 0; s < NSQ; s++)
    if (b[s] == mk(T_KING, color)) return s;
  return -1;
}

/* check_if_square_is_attacked: shogi_rules_logic.py:234-272 */
static int square_attacked(const int8_t *b, int target, int attacker_color) {
  int tg[40];
  for (int s = 0; s < NSQ; s++) {
    int code = b[s];
    if (code && pcolor(code) == attacker_color) {
      int n = potential_moves(b, code, s / 9, s % 9, tg);
      for (int i = 0; i < n; i++)
        if (tg[i] == target) return 1;
    }
  }
  return 0;
}

/* is_in_check / is_king_in_check_after_simulated_move:
 * shogi_rules_logic.py:36-67, 362-379 -- a missing king counts as "in check". */
static int king_in_check(const int8_t *b, int color) {
  int k = find_king(b, color);
  if (k < 0) return 1;
  return square_attacked(b, k, 1 - color);
}

/* ShogiGame.is_in_promotion_zone: shogi_game.py:821-825 */
static inline int in_zone(int row, int color) { return color == 0 ? (row <= 2) : (row >= 6); }
/* can_promote_specific_piece: shogi_rules_logic.py:382-401 */
static int can_promote(int code, int r_from, int r_to) {
  int t = ptype(code);
  if (t == T_GOLD || t == T_KING || t >= T_PPAWN) return 0;
  return in_zone(r_from, pcolor(code)) || in_zone(r_to, pcolor(code));
}
/* must_promote_specific_piece: shogi_rules_logic.py:404-421 */
static int must_promote(int code, int r_to) {
  int t = ptype(code), c = pcolor(code);
  if (t == T_PAWN || t == T_LANCE) return (c == 0 && r_to == 0) || (c == 1 && r_to == 8);
  if (t == T_KNIGHT) return (c == 0 && r_to <= 1) || (c == 1 && r_to >= 7);
  return 0;
}

/* check_for_nifu: shogi_rules_logic.py:211-231 */
static int nifu(const int8_t *b, int color, int col) {
  for (int r = 0; r < 9; r++)
    if (b[r * 9 + col] == mk(T_PAWN, color)) return 1;
  return 0;
}

/* PolicyOutputMapper closed form: keisei/utils/utils.py:208-266 */
static inline int board_move_index(int from, int to, int promo) {
  return ((from * 80 + to - (to > from)) * 2 + promo);
}
static inline int drop_index(int to, int t) { return 12960 + to * 7 + t; }

/* apply_move_to_board_state: shogi_move_execution.py:24-138 (board + hands only) */
static void apply_board_move(int8_t *b, int32_t hands[2][8], int from, int to, int promo, int mover) {
  int moving = b[from];
  int target = b[to];
  if (target) hands[mover][base_of(ptype(target))] += 1; /* :108-114 (incl. the KING quirk) */
  b[to] = (int8_t)moving;
  b[from] = 0;
  if (promo) b[to] = (int8_t)mk(promoted_of(ptype(moving)), pcolor(moving)); /* :122-132 */
}
static void apply_drop(int8_t *b, int32_t hands[2][8], int to, int t, int mover) {
  b[to] = (int8_t)mk(t, mover); /* :64 */
  hands[mover][t] -= 1;         /* :71 */
}

static int gen_legal(orc_game *g, int escape_mode, uint16_t *out);

/* check_for_uchi_fu_zume: shogi_rules_logic.py:275-359 */
static int uchi_fu_zume(orc_game *g, int sq, int color) {
  int opp = 1 - color;
  if (g->board[sq] != 0) return 0;          /* :299-301 */
  if (g->hands[color][T_PAWN] <= 0) return 0; /* :304-309 */
  int saved_side = g->side;
  g->board[sq] = (int8_t)mk(T_PAWN, color); /* :313-314 */
  g->hands[color][T_PAWN] -= 1;
  int result = 0;
  int k = find_king(g->board, opp);
  if (k >= 0 && square_attacked(g->board, k, color)) { /* :321-341 */
    g->side = opp;                                     /* :345 */
    uint16_t tmp[1024];
    int n = gen_legal(g, 1, tmp);                      /* :347 */
    result = (n == 0);                                 /* :357 */
  }
  g->board[sq] = 0;                                    /* :351-353 */
  g->hands[color][T_PAWN] += 1;
  g->side = saved_side;
  return result;
}

/* can_drop_specific_piece: shogi_rules_logic.py:424-483 */
static int can_drop(orc_game *g, int t, int sq, int color, int escape_mode) {
  if (g->board[sq] != 0) return 0;
  int r = sq / 9, c = sq % 9;
  int last = color == 0 ? 0 : 8, second = color == 0 ? 1 : 7;
  if (t == T_PAWN) {
    if (nifu(g->board, color, c)) return 0;
    if (r == last) return 0;
    if (!escape_mode && uchi_fu_zume(g, sq, color)) return 0;
  } else if (t == T_LANCE) {
    if (r == last) return 0;
  } else if (t == T_KNIGHT) {
    if (r == last || r == second) return 0;
  }
  return 1;
}

/* generate_all_legal_moves: shogi_rules_logic.py:486-635.
 * Every candidate is applied to a copy (the reference's make_move(is_simulation=True)
 * deep copy, shogi_game.py:634-636) and kept iff the mover's king exists and is not
 * attacked afterwards (:550-554, :616-622).  Output = policy indices in the
 * reference's generation order (only the SET is contractual). */
static int gen_legal(orc_game *g, int escape_mode, uint16_t *out) {
  int n = 0;
  int me = g->side;
  int8_t sim[NSQ];
  int32_t simh[2][8];
  int tg[40];
  for (int from = 0; from < NSQ; from++) {
    int code = g->board[from];
    if (!code || pcolor(code) != me) continue;
    int nt = potential_moves(g->board, code, from / 9, from % 9, tg);
    for (int i = 0; i < nt; i++) {
      int to = tg[i];
      int cp = can_promote(code, from / 9, to / 9);
      int mp = must_promote(code, to / 9);
      for (int promo = 0; promo <= 1; promo++) { /* :511-519 */
        if (promo == 0 && mp) continue;
        if (promo == 1 && !(cp || mp)) continue;
        memcpy(sim, g->board, NSQ);
        memcpy(simh, g->hands, sizeof simh);
        apply_board_move(sim, simh, from, to, promo, me);
        if (!king_in_check(sim, me)) out[n++] = (uint16_t)board_move_index(from, to, promo);
      }
    }
  }
  for (int t = 0; t < 7; t++) { /* hand-dict order P,L,N,S,G,B,R: :562-566 */
    if (g->hands[me][t] <= 0) continue;
    for (int sq = 0; sq < NSQ; sq++) {
      if (!can_drop(g, t, sq, me, escape_mode)) continue;
      memcpy(sim, g->board, NSQ);
      memcpy(simh, g->hands, sizeof simh);
      apply_drop(sim, simh, sq, t, me);
      if (!king_in_check(sim, me)) out[n++] = (uint16_t)drop_index(sq, t);
    }
  }
  return n;
}

/* _board_state_hash: shogi_game.py:347-372 -- the full state, not a hash.  Equality of the
 * "sorted non-zero (type,count)" tuples is equality of the count vectors. */
static void snapshot(const orc_game *g, uint8_t *dst) {
  memcpy(dst, g->board, NSQ);
  for (int c = 0; c < 2; c++)
    for (int t = 0; t < 8; t++) dst[NSQ + c * 8 + t] = (uint8_t)g->hands[c][t];
  dst[NSQ + 16] = (uint8_t)g->side;
}

/* check_for_sennichite: shogi_rules_logic.py:638-695 */
static int sennichite(const orc_game *g) {
  if (g->hist_len == 0) return 0;
  const uint8_t *last = g->hist + (size_t)(g->hist_len - 1) * SNAP;
  int count = 0;
  for (int i = 0; i < g->hist_len; i++)
    if (memcmp(g->hist + (size_t)i * SNAP, last, SNAP) == 0) count++;
  return count >= 4;
}

/* _check_and_update_termination_status: shogi_game.py:408-450 */
static void update_termination(orc_game *g, int mover) {
  if (g->game_over) return;
  uint16_t tmp[1024];
  int n = gen_legal(g, 0, tmp);
  if (n == 0) {
    if (king_in_check(g->board, g->side)) { g->game_over = 1; g->winner = mover; g->reason = R_TSUMI; }
    else { g->game_over = 1; g->winner = -1; g->reason = R_STALEMATE; }
    return;
  }
  if (g->move_count >= g->max_moves) { g->game_over = 1; g->winner = -1; g->reason = R_MAXMOVES; return; }
  if (sennichite(g)) { g->game_over = 1; g->winner = -1; g->reason = R_SENNICHITE; }
}

/* ------------------------------------------------------------------ public API */

orc_game *orc_new(void) {
  orc_game *g = (orc_game *)calloc(1, sizeof(orc_game));
  return g;
}
void orc_free(orc_game *g) {
  if (g) { free(g->hist); free(g); }
}

/* ShogiGame.reset / _setup_initial_board: shogi_game.py:79-130 */
void orc_reset(orc_game *g, int max_moves) {
  static const int back[9] = {T_LANCE, T_KNIGHT, T_SILVER, T_GOLD, T_KING, T_GOLD, T_SILVER, T_KNIGHT, T_LANCE};
  memset(g->board, 0, NSQ);
  for (int c = 0; c < 9; c++) {
    g->board[0 * 9 + c] = (int8_t)mk(back[c], 1);
    g->board[8 * 9 + c] = (int8_t)mk(back[c], 0);
    g->board[2 * 9 + c] = (int8_t)mk(T_PAWN, 1);
    g->board[6 * 9 + c] = (int8_t)mk(T_PAWN, 0);
  }
  g->board[1 * 9 + 1] = (int8_t)mk(T_ROOK, 1);
  g->board[1 * 9 + 7] = (int8_t)mk(T_BISHOP, 1);
  g->board[7 * 9 + 1] = (int8_t)mk(T_BISHOP, 0);
  g->board[7 * 9 + 7] = (int8_t)mk(T_ROOK, 0);
  memset(g->hands, 0, sizeof g->hands);
  g->side = 0;
  g->move_count = 0;
  g->max_moves = max_moves;
  g->game_over = 0;
  g->winner = -1;
  g->reason = R_NONE;
  g->hist_len = 0;
}

/* ShogiGame.from_sfen after parsing: shogi_game.py:306-345.  hands14 = black P..R then white P..R.
 * evaluate_termination mirrors the call at :343 (mover := opponent of side to move). */
void orc_load(orc_game *g, const int8_t *board81, const uint8_t *hands14, int side, int move_count,
              int max_moves, int evaluate_termination) {
  memcpy(g->board, board81, NSQ);
  memset(g->hands, 0, sizeof g->hands);
  for (int c = 0; c < 2; c++)
    for (int t = 0; t < 7; t++) g->hands[c][t] = hands14[c * 7 + t];
  g->side = side;
  g->move_count = move_count;
  g->max_moves = max_moves;
  g->game_over = 0;
  g->winner = -1;
  g->reason = R_NONE;
  g->hist_len = 0;
  if (evaluate_termination) update_termination(g, 1 - side);
}

void orc_export(const orc_game *g, int8_t *board81, uint8_t *hands14, int32_t *meta6) {
  memcpy(board81, g->board, NSQ);
  for (int c = 0; c < 2; c++)
    for (int t = 0; t < 7; t++) hands14[c * 7 + t] = (uint8_t)g->hands[c][t];
  meta6[0] = g->side; meta6[1] = g->move_count; meta6[2] = g->max_moves;
  meta6[3] = g->game_over; meta6[4] = g->winner; meta6[5] = g->reason;
}

/* ShogiGame.get_legal_moves: shogi_game.py:216-218.  Returns count; out (>=1024 entries) gets
 * policy indices.  The reference's simulation undo clears game_over/winner/reason
 * (shogi_move_execution.py:218-221) whenever at least one candidate was simulated; the
 * scalar facade mirrors that, the oracle exposes it through *simulated. */
int orc_legal_moves(orc_game *g, uint16_t *out) { return gen_legal(g, 0, out); }

int orc_legal_mask(orc_game *g, uint8_t *mask) { /* utils.py:310-336 */
  uint16_t tmp[1024];
  int n = gen_legal(g, 0, tmp);
  memset(mask, 0, NACT);
  for (int i = 0; i < n; i++) mask[tmp[i]] = 1;
  return n;
}

int orc_in_check(const orc_game *g, int color) { return king_in_check(g->board, color); }

/* generate_piece_potential_moves (shogi_rules_logic.py:82-208) of the piece on sq: out81[t] = 1 for every target */
int orc_piece_targets(const orc_game *g, int sq, uint8_t *out81) {
  memset(out81, 0, NSQ);
  if (sq < 0 || sq >= NSQ || g->board[sq] == 0) return 0;
  int t[64];
  int n = potential_moves(g->board, g->board[sq], sq / 9, sq % 9, t);
  for (int i = 0; i < n; i++) out81[t[i]] = 1;
  return n;
}

/* ShogiGame.is_uchi_fu_zume (shogi_game.py:237-241 -> check_for_uchi_fu_zume, shogi_rules_logic.py:275-359) */
int orc_uchi_fu_zume(orc_game *g, int sq, int color) { return uchi_fu_zume(g, sq, color); }

/* ShogiGame.can_drop_piece (shogi_game.py:243-260 -> can_drop_specific_piece, shogi_rules_logic.py:424-483), after the
 * facade's own guards: piece in hand, square on the board */
int orc_can_drop(orc_game *g, int type, int sq, int color) {
  if (type < 0 || type > 6 || sq < 0 || sq >= NSQ) return 0;
  if (g->hands[color][type] <= 0) return 0;
  return can_drop(g, type, sq, color, 0);
}

/* ShogiGame.get_king_legal_moves (shogi_game.py:208-235): legal moves of `color` whose mover is the king */
int orc_king_legal_moves(orc_game *g, int color) {
  if (find_king(g->board, color) < 0) return 0;
  int saved = g->side;
  g->side = color;
  uint16_t tmp[1024];
  int n = gen_legal(g, 0, tmp), k = 0;
  for (int i = 0; i < n; i++) {
    if (tmp[i] >= 12960) continue;
    int from = (tmp[i] >> 1) / 80;
    if (ptype(g->board[from]) == T_KING) k++;
  }
  g->side = saved;
  return k;
}

/* generate_neural_network_observation: shogi_game_io.py:434-539 */
void orc_observation(const orc_game *g, float *obs) {
  memset(obs, 0, sizeof(float) * 46 * NSQ);
  int me = g->side;
  for (int s = 0; s < NSQ; s++) {
    int code = g->board[s];
    if (!code) continue;
    int fs = me == 0 ? s : 80 - s; /* (8-r, 8-c): :468-469 */
    int t = ptype(code), mine = pcolor(code) == me;
    int ch;
    if (t >= T_PPAWN) ch = (mine ? 8 : 22) + (t - T_PPAWN); /* :478-486 */
    else ch = (mine ? 0 : 14) + t;                          /* :487-495 */
    obs[ch * NSQ + fs] = 1.0f;
  }
  for (int t = 0; t < 7; t++) { /* :506-527 */
    int a = g->hands[me][t], b = g->hands[1 - me][t];
    if (a > 0) { float v = (float)((double)a / 18.0); for (int s = 0; s < NSQ; s++) obs[(28 + t) * NSQ + s] = v; }
    if (b > 0) { float v = (float)((double)b / 18.0); for (int s = 0; s < NSQ; s++) obs[(35 + t) * NSQ + s] = v; }
  }
  float side_v = me == 0 ? 1.0f : 0.0f; /* :530-532 */
  float mc = g->max_moves > 0 ? (float)((double)g->move_count / (double)g->max_moves) : 0.0f; /* :535-536 */
  for (int s = 0; s < NSQ; s++) { obs[42 * NSQ + s] = side_v; obs[43 * NSQ + s] = mc; }
}

/* ShogiGame.make_move (real move): shogi_game.py:574-660.
 * Returns 0 ok, -1 malformed index, -2 no piece / wrong colour, -3 illegal movement pattern
 * (the ValueError cases of :461-483, :529-544).  out4 = {reward, done, reason, winner}. */
int orc_make_move(orc_game *g, int action, float *out4) {
  int mover;
  if (g->game_over) { /* :589-593 */
    mover = 1 - g->side;
  } else {
    if (action < 0 || action >= NACT) return -1;
    mover = g->side;
    if (action < 12960) {
      int promo = action & 1, pair = action >> 1;
      int from = pair / 80, t = pair % 80, to = t + (t >= from);
      int code = g->board[from];
      if (!code || pcolor(code) != mover) return -2;
      int tg[40];
      int nt = potential_moves(g->board, code, from / 9, from % 9, tg);
      int ok = 0;
      for (int i = 0; i < nt; i++) ok |= (tg[i] == to);
      if (!ok) return -3;
      apply_board_move(g->board, g->hands, from, to, promo, mover);
    } else {
      int k = action - 12960;
      apply_drop(g->board, g->hands, k / 7, k % 7, mover);
    }
    g->move_count += 1;  /* apply_move_to_game: shogi_move_execution.py:141-156 */
    g->side = 1 - mover;
    if (g->hist_len == g->hist_cap) { /* :651-654 */
      g->hist_cap = g->hist_cap ? g->hist_cap * 2 : 64;
      g->hist = (uint8_t *)realloc(g->hist, (size_t)g->hist_cap * SNAP);
    }
    snapshot(g, g->hist + (size_t)g->hist_len * SNAP);
    g->hist_len++;
    update_termination(g, mover); /* :655 */
  }
  float reward = 0.0f; /* _handle_real_move_return: :553-572 */
  if (g->game_over && g->winner >= 0) reward = g->winner == mover ? 1.0f : -1.0f;
  out4[0] = reward;
  out4[1] = (float)g->game_over;
  out4[2] = (float)g->reason;
  out4[3] = (float)g->winner;
  return 0;
}

/* ---- counter-based action RNG shared (by specification, not by code) with the CUDA path:
 * r = hi32(splitmix64-finaliser(seed ^ env*0x9E3779B97F4A7C15 ^ step*0xBF58476D1CE4E5B9)),
 * k = (r * n_legal) >> 32, action = k-th legal index in ascending policy-index order. */
uint32_t orc_rand32(uint64_t seed, uint64_t env, uint64_t step) {
  uint64_t x = seed ^ (env * 0x9E3779B97F4A7C15ull) ^ (step * 0xBF58476D1CE4E5B9ull);
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (uint32_t)(x >> 32);
}

static int cmp_u16(const void *a, const void *b) { return (int)*(const uint16_t *)a - (int)*(const uint16_t *)b; }

int orc_pick_action(orc_game *g, uint64_t seed, uint64_t env, uint64_t step) {
  uint16_t tmp[1024];
  int n = gen_legal(g, 0, tmp);
  if (n == 0) return -1;
  qsort(tmp, n, sizeof(uint16_t), cmp_u16);
  uint32_t r = orc_rand32(seed, env, step);
  return tmp[((uint64_t)r * (uint64_t)n) >> 32];
}

/* Random-legal self-play of n_envs independent games for T steps with auto-reset, the CPU
 * mirror of BASELINE config 2 (and the loop BASELINE.md section 3.1 times).  Per step it does what
 * StepManager.execute_step does around the engine: get_legal_moves -> get_legal_mask ->
 * make_move -> observation (step_manager.py:117,154,229).  Optional outputs (NULL to skip):
 *   actions/rewards/dones/reasons/legal_counts [T*n] (legal count of the state the action was
 *   chosen in), obs_out [n,46*81] / mask_out [n,13527] / boards_out [n,81] / hands_out [n,14] /
 *   meta_out [n,6] = the state after the LAST step.  Returns total steps executed.
 *   Envs are spread over n_threads POSIX threads (dynamic, one env at a time). */
typedef struct {
  int n_envs, env0, T, step0, max_moves;
  uint64_t seed;
  int32_t *actions; float *rewards; uint8_t *dones; uint8_t *reasons; int32_t *legal_counts;
  float *obs_out; uint8_t *mask_out; int8_t *boards_out; uint8_t *hands_out; int32_t *meta_out;
  const int8_t *boards0; const uint8_t *hands0; const uint8_t *sides0; const int32_t *mc0; /* optional start positions */
  int next_env;
  long total;
  pthread_mutex_t mu;
} sp_job;

static void *sp_worker(void *arg) {
  sp_job *j = (sp_job *)arg;
  orc_game *g = orc_new();
  float *obs = (float *)malloc(sizeof(float) * 46 * NSQ);
  uint8_t *mask = (uint8_t *)malloc(NACT);
  uint16_t tmp[1024];
  long mine = 0;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    int e = j->next_env++;
    pthread_mutex_unlock(&j->mu);
    if (e >= j->n_envs) break;
    orc_reset(g, j->max_moves);
    if (j->boards0)
      orc_load(g, j->boards0 + (size_t)e * NSQ, j->hands0 + (size_t)e * 14, j->sides0[e], j->mc0[e], j->max_moves, 0);
    for (int t = 0; t < j->T; t++) {
      int n = gen_legal(g, 0, tmp);
      memset(mask, 0, NACT);
      for (int i = 0; i < n; i++) mask[tmp[i]] = 1;
      qsort(tmp, n, sizeof(uint16_t), cmp_u16);
      uint32_t r = orc_rand32(j->seed, (uint64_t)(j->env0 + e), (uint64_t)(j->step0 + t));
      int a = n ? tmp[((uint64_t)r * (uint64_t)n) >> 32] : -1;
      if (n == 0) { /* loaded position without a legal move: the caller filters these out */
        size_t ix0 = (size_t)t * j->n_envs + e;
        if (j->actions) j->actions[ix0] = -1;
        if (j->legal_counts) j->legal_counts[ix0] = 0;
        continue;
      }
      float o4[4];
      orc_make_move(g, a, o4);
      orc_observation(g, obs);
      size_t ix = (size_t)t * j->n_envs + e;
      if (j->actions) j->actions[ix] = a;
      if (j->rewards) j->rewards[ix] = o4[0];
      if (j->dones) j->dones[ix] = (uint8_t)o4[1];
      if (j->reasons) j->reasons[ix] = (uint8_t)o4[2];
      if (j->legal_counts) j->legal_counts[ix] = n;
      if (g->game_over) orc_reset(g, j->max_moves); /* handle_episode_end: step_manager.py:437-440 */
      mine++;
    }
    if (j->obs_out) orc_observation(g, j->obs_out + (size_t)e * 46 * NSQ);
    if (j->mask_out) orc_legal_mask(g, j->mask_out + (size_t)e * NACT);
    if (j->boards_out) {
      int32_t m6[6];
      orc_export(g, j->boards_out + (size_t)e * NSQ, j->hands_out + (size_t)e * 14, m6);
      if (j->meta_out) memcpy(j->meta_out + (size_t)e * 6, m6, sizeof m6);
    }
  }
  free(obs); free(mask);
  orc_free(g);
  pthread_mutex_lock(&j->mu);
  j->total += mine;
  pthread_mutex_unlock(&j->mu);
  return NULL;
}

long orc_selfplay(int n_envs, int env0, int T, int step0, int max_moves, uint64_t seed, int32_t *actions,
                  float *rewards, uint8_t *dones, uint8_t *reasons, int32_t *legal_counts,
                  float *obs_out, uint8_t *mask_out, int8_t *boards_out, uint8_t *hands_out,
                  int32_t *meta_out, int n_threads, const int8_t *boards0, const uint8_t *hands0,
                  const uint8_t *sides0, const int32_t *mc0) {
  sp_job j = {n_envs, env0, T, step0, max_moves, seed, actions, rewards, dones, reasons, legal_counts,
              obs_out, mask_out, boards_out, hands_out, meta_out, boards0, hands0, sides0, mc0,
              0, 0, PTHREAD_MUTEX_INITIALIZER};
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, sp_worker, &j);
  for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
  return j.total;
}

/* Persistent batch of games for stepwise timing (bench.py --impl reference / cpu_baseline):
 * orc_batch_run advances every game of the batch by T plies with the same per-ply work as
 * orc_selfplay (legal moves -> mask -> make_move -> observation -> reset on done). */
typedef struct {
  int n, max_moves;
  orc_game **games;
} orc_batch;

orc_batch *orc_batch_new(int n, int max_moves) {
  orc_batch *b = (orc_batch *)calloc(1, sizeof(orc_batch));
  b->n = n; b->max_moves = max_moves;
  b->games = (orc_game **)calloc((size_t)n, sizeof(orc_game *));
  for (int i = 0; i < n; i++) { b->games[i] = orc_new(); orc_reset(b->games[i], max_moves); }
  return b;
}
void orc_batch_free(orc_batch *b) {
  if (!b) return;
  for (int i = 0; i < b->n; i++) orc_free(b->games[i]);
  free(b->games); free(b);
}
typedef struct {
  orc_batch *b; int T, step0, env0; uint64_t seed; int next_env; long total; long ply_sum; pthread_mutex_t mu;
} bt_job;
static void *bt_worker(void *arg) {
  bt_job *j = (bt_job *)arg;
  float *obs = (float *)malloc(sizeof(float) * 46 * NSQ);
  uint8_t *mask = (uint8_t *)malloc(NACT);
  uint16_t tmp[1024];
  long mine = 0, plies = 0;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    int e = j->next_env++;
    pthread_mutex_unlock(&j->mu);
    if (e >= j->b->n) break;
    orc_game *g = j->b->games[e];
    for (int t = 0; t < j->T; t++) {
      int n = gen_legal(g, 0, tmp);
      memset(mask, 0, NACT);
      for (int i = 0; i < n; i++) mask[tmp[i]] = 1;
      qsort(tmp, n, sizeof(uint16_t), cmp_u16);
      uint32_t r = orc_rand32(j->seed, (uint64_t)(j->env0 + e), (uint64_t)(j->step0 + t));
      float o4[4];
      plies += g->move_count;
      orc_make_move(g, tmp[((uint64_t)r * (uint64_t)n) >> 32], o4);
      orc_observation(g, obs);
      if (g->game_over) orc_reset(g, j->b->max_moves);
      mine++;
    }
  }
  free(obs); free(mask);
  pthread_mutex_lock(&j->mu);
  j->total += mine; j->ply_sum += plies;
  pthread_mutex_unlock(&j->mu);
  return NULL;
}
/* returns env steps executed; *mean_ply (optional) = mean move_count of the positions stepped from */
long orc_batch_run(orc_batch *b, int T, int step0, int env0, uint64_t seed, int n_threads, double *mean_ply) {
  bt_job j = {b, T, step0, env0, seed, 0, 0, 0, PTHREAD_MUTEX_INITIALIZER};
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, bt_worker, &j);
  for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
  if (mean_ply) *mean_ply = j.total ? (double)j.ply_sum / (double)j.total : 0.0;
  return j.total;
}

/* ExperienceBuffer.compute_advantages_and_returns: keisei/core/experience_buffer.py:99-145,
 * column-wise over a [T, N] layout (N = 1 is the reference's flat buffer).  Every fp32
 * operation is rounded separately, in the reference's order (build with -ffp-contract=off):
 *   delta = (r + (f32(gamma) * nv) * m) - V ;  gae = delta + (f32(gamma*lambda) * m) * gae
 * with gamma*lambda multiplied in double first (:138). */
void orc_gae(const float *rewards, const float *values, const uint8_t *dones, const float *last_value,
             int T, int N, double gamma, double lambda, float *adv, float *ret) {
  volatile float g32 = (float)gamma;
  volatile float gl32 = (float)(gamma * lambda);
  for (int n = 0; n < N; n++) {
    float gae = 0.0f;
    for (int t = T - 1; t >= 0; t--) {
      size_t i = (size_t)t * N + n;
      float m = 1.0f - (dones[i] ? 1.0f : 0.0f);
      float nv = (t == T - 1) ? last_value[n] : values[i + N];
      volatile float a = g32 * nv;
      volatile float b = a * m;
      volatile float c = rewards[i] + b;
      float delta = c - values[i];
      volatile float d = gl32 * m;
      volatile float e = d * gae;
      gae = delta + e;
      adv[i] = gae;
      ret[i] = gae + values[i];
    }
  }
}
