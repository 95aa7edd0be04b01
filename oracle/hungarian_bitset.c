/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * CPU model of the duplicate-free reformulation that the CUDA kernel
 * (rec-attend-public_b200/csrc/hungarian.cu) uses for the reference's
 * breadth-first augmenting-path search (/root/reference/hungarian.cc:107-217).
 * It exists so that the equivalence "reformulated search == literal search" can
 * be hammered with millions of random and degenerate inputs on the CPU
 * (tests/test_hungarian_oracle.py) instead of on scarce GPU time.
 *
 * Why the reformulation is exact.  The reference BFS marks a vertex when it is
 * popped, so the FIFO holds many copies of a vertex and the parent of a vertex
 * is the LAST vertex that pushed it.  In the source -> X -> Y -> sink network all
 * edges go between consecutive BFS levels, therefore
 *   (1) every copy of a vertex of level L sits in the queue segment of level L
 *       and is pushed only while segment L-1 is being popped;
 *   (2) parent(u) = the neighbour v of level L-1 whose LAST copy is popped
 *       latest, i.e. the neighbour of highest "rank";
 *   (3) the rank order inside level L is (rank of parent, vertex index),
 *       because a popped copy pushes its unmarked neighbours in index order;
 *   (4) level 1 is the unmatched rows in index order; an X vertex of a deeper
 *       level is reached only from the column it is matched to;
 *   (5) the sink's parent is the highest-ranked unmatched column of the first
 *       Y level that holds an unmatched column.
 * Ranks are all that is needed to reproduce the parent chain, hence the
 * augmenting path, hence the matching, bit for bit — without the duplicate
 * copies (and therefore without the reference's 1000-pop abort).
 *
 * The cover/equality-graph outer loop (hungarian.cc:335-488) is kept as is,
 * with the vertex sets S, T, N_S held as 64-bit masks (nx, ny <= 64).
 */
#include <float.h>
#include <stdint.h>
#include <string.h>

#define RA_HUNG_EPS 1e-6
#define RA_HUNG_MAX_ITER 1000
#define RA_ST_OUTER_CAP 1
#define RA_ST_TOO_LARGE 16

typedef uint64_t mask_t;

static inline int lowest_bit(mask_t m) { return __builtin_ctzll(m); }

/* One search + augmentation over the equality graph `eq` (row masks).
 * match_y[x] / match_x[y] hold the current matching (-1 = free). Returns 1 if augmented. */
static int augment_ranked(const mask_t *eq, int nx, int ny, int *match_y, int *match_x) {
  int level_x[64]; /* X vertices of the current level, ascending rank */
  int level_y[64];
  int parent_of_y[64];
  int n_lx = 0;
  mask_t seen_y = 0;
  (void)ny;

  for (int x = 0; x < nx; ++x)
    if (match_y[x] < 0) level_x[n_lx++] = x;

  while (n_lx > 0) {
    /* (2): highest-ranked pusher wins, so claim columns from the back of the level */
    mask_t claimed[64];
    mask_t level_claim = 0;
    for (int r = n_lx - 1; r >= 0; --r) {
      int x = level_x[r];
      mask_t adj = eq[x];
      if (match_y[x] >= 0) adj &= ~((mask_t)1 << match_y[x]); /* saturated edge has no residual */
      mask_t c = adj & ~seen_y & ~level_claim;
      claimed[r] = c;
      level_claim |= c;
    }
    if (!level_claim) return 0;
    seen_y |= level_claim;

    /* (3): new level in (parent rank, index) order */
    int n_ly = 0;
    int last_free_y = -1;
    for (int r = 0; r < n_lx; ++r) {
      mask_t c = claimed[r];
      while (c) {
        int y = lowest_bit(c);
        c &= c - 1;
        parent_of_y[y] = level_x[r];
        level_y[n_ly++] = y;
        if (match_x[y] < 0) last_free_y = y; /* (5) */
      }
    }

    if (last_free_y >= 0) {
      int y = last_free_y;
      for (;;) {
        int x = parent_of_y[y];
        int prev = match_y[x];
        match_y[x] = y;
        match_x[y] = x;
        if (prev < 0) break;
        y = prev;
      }
      return 1;
    }

    /* (4): every column of the level is matched; step back along matched edges */
    n_lx = 0;
    for (int r = 0; r < n_ly; ++r) level_x[n_lx++] = match_x[level_y[r]];
  }
  return 0;
}

static int cover_one(const float *w, int nx, int ny, float *M, float *cx, float *cy) {
  mask_t eq[64];
  int match_y[64], match_x[64];
  mask_t S = 0, T = 0;
  int status = 0;
  const mask_t all_y = ny == 64 ? ~(mask_t)0 : (((mask_t)1 << ny) - 1);

  for (int x = 0; x < nx; ++x) {
    float mx = w[x * ny];
    for (int y = 1; y < ny; ++y)
      if (w[x * ny + y] > mx) mx = w[x * ny + y];
    cx[x] = mx;
    match_y[x] = -1;
  }
  for (int y = 0; y < ny; ++y) {
    cy[y] = 0.0f;
    match_x[y] = -1;
  }

  int next_match = 1;
  for (long round = 0;; ++round) {
    if (round == RA_HUNG_MAX_ITER) {
      status |= RA_ST_OUTER_CAP;
      break;
    }
    for (int x = 0; x < nx; ++x) {
      mask_t row = 0;
      for (int y = 0; y < ny; ++y) {
        float d = cx[x] + cy[y] - w[x * ny + y];
        float ad = d > 0 ? d : -d;
        if ((double)ad <= RA_HUNG_EPS && (cx[x] > 0 || cy[y] > 0)) row |= (mask_t)1 << y;
      }
      eq[x] = row;
    }
    if (next_match) {
      for (int x = 0; x < nx; ++x) match_y[x] = -1;
      for (int y = 0; y < ny; ++y) match_x[y] = -1;
      while (augment_ranked(eq, nx, ny, match_y, match_x)) {
      }
      int n_matched = 0, first_free = -1;
      for (int x = 0; x < nx; ++x) {
        if (match_y[x] >= 0)
          ++n_matched;
        else if (first_free < 0)
          first_free = x;
      }
      /* hungarian.cc:219-248: the smaller side must be fully matched */
      if (n_matched == (nx >= ny ? ny : nx)) break;
      S = (mask_t)1 << first_free;
      T = 0;
    }

    mask_t NS = 0;
    for (int x = 0; x < nx; ++x)
      if ((S >> x) & 1) NS |= eq[x];

    if (NS == T) {
      float a = FLT_MAX;
      for (int x = 0; x < nx; ++x)
        if ((S >> x) & 1)
          for (int y = 0; y < ny; ++y)
            if (!((T >> y) & 1)) {
              float d = cx[x] + cy[y] - w[x * ny + y];
              if (d < a) a = d;
            }
      if ((double)a < RA_HUNG_EPS) {
        next_match = 1;
        continue;
      }
      for (int x = 0; x < nx; ++x)
        if ((S >> x) & 1) cx[x] -= a;
      for (int y = 0; y < ny; ++y)
        if ((T >> y) & 1) cy[y] += a;
    } else {
      while (__builtin_popcountll(NS) > __builtin_popcountll(T)) {
        int y = lowest_bit(NS & ~T & all_y);
        int z = match_x[y];
        if (z < 0) {
          next_match = 1;
          break;
        }
        next_match = 0;
        S |= (mask_t)1 << z;
        NS |= eq[z];
        T |= (mask_t)1 << y;
      }
    }
  }

  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y) M[x * ny + y] = (match_y[x] == y) ? 1.0f : 0.0f;
  return status;
}

int ra_oracle_hungarian_bitset_f32(const float *W, int B, int nx, int ny, float *M, float *cx, float *cy, int *status) {
  int all = 0;
  if (nx < 1 || ny < 1 || nx > 64 || ny > 64) {
    for (int b = 0; b < B && status; ++b) status[b] = RA_ST_TOO_LARGE;
    return RA_ST_TOO_LARGE;
  }
  for (int b = 0; b < B; ++b) {
    int st = cover_one(W + (size_t)b * nx * ny, nx, ny, M + (size_t)b * nx * ny, cx + (size_t)b * nx,
                       cy + (size_t)b * ny);
    if (status) status[b] = st;
    all |= st;
  }
  return all;
}
