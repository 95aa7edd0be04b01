/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C restatement of the reference's `Hungarian` TensorFlow custom op
 * (/root/reference/hungarian.cc), re-expressing its algorithm function by
 * function with plain arrays (no TensorFlow, no Eigen).  It is PINNED twice:
 * against the reference's own known-answer tests
 * (/root/reference/hungarian_tf_tests.py:9-91, fixtures in
 * tests/golden/hungarian_kat.json), and against the reference's own
 * hungarian.cc compiled unmodified over stand-in TensorFlow / Eigen headers
 * (oracle/_ref/libhungarian_ref.so, see oracle/Makefile) - bit-identical on
 * 20k random problems (tests/test_hungarian_oracle.py).
 *
 * Everything that decides tie-breaking is kept literal:
 *   - fp32 cover arithmetic, `|cx+cy-w| <= 1e-6` compared in double
 *     (hungarian.cc:18,318-319), and the `(cx>0 || cy>0)` guard;
 *   - the breadth-first search marks vertices when they are POPPED, pushes a
 *     vertex again every time an unmarked copy is seen and overwrites its
 *     parent on every push (hungarian.cc:124-141);
 *   - the matching is rebuilt from zero flow on every `next_match` round
 *     (hungarian.cc:179-217);
 *   - sets S, T, N_S are iterated in ascending vertex order (std::set);
 *   - the first unmatched row seeds S (hungarian.cc:394-403);
 *   - iteration caps: the reference aborts the process (LOG(FATAL)) when a BFS
 *     pops 1000 vertices or a max-flow runs 1000 augmentations, and returns the
 *     unfinished matching (LOG(ERROR)) after 1000 outer rounds
 *     (hungarian.cc:20,124-127,184-188,362-377).  Here the fatal caps become
 *     status bits; with `stop_at_fatal == 0` the search simply continues
 *     (the "uncapped" algorithm the CUDA kernel implements), with
 *     `stop_at_fatal != 0` the example is abandoned like the reference would.
 */
#include <float.h>
#include <stdlib.h>
#include <string.h>

#define RA_HUNG_EPS 1e-6 /* double literal on purpose: hungarian.cc:18 */
#define RA_HUNG_MAX_ITER 1000

#define RA_ST_OUTER_CAP 1   /* outer loop hit 1000 rounds: partial matching returned */
#define RA_ST_BFS_CAP 2     /* a BFS popped 1000 vertices: reference would LOG(FATAL) */
#define RA_ST_FLOW_CAP 4    /* a max-flow ran 1000 augmentations: reference would LOG(FATAL) */
#define RA_ST_EQUALIZE_CAP 8 /* N_S/T equalising loop ran 1000 times: LOG(FATAL) */

typedef struct {
  int n;            /* vertices of the flow network: source, X, Y, sink */
  float *capacity;  /* n*n */
  float *flow;      /* n*n */
  float *residual;  /* n*n */
  unsigned char *mark;
  int *parent;
  int *queue;
  size_t queue_cap;
  long max_pops; /* largest pop count seen by any BFS (diagnostic) */
  int status;
  int stop_at_fatal;
} flow_net;

static void queue_push(flow_net *g, size_t *tail, int v) {
  if (*tail == g->queue_cap) {
    g->queue_cap *= 2;
    g->queue = (int *)realloc(g->queue, g->queue_cap * sizeof(int));
  }
  g->queue[(*tail)++] = v;
}

/* hungarian.cc:107-177 (Augment): one BFS from the source, then push one unit
 * of flow along the parent chain if the sink was reached. */
static int augment(flow_net *g) {
  const int n = g->n;
  const int s = 0, t = n - 1;
  size_t head = 0, tail = 0;
  int found = 0;
  long pops = 0;

  memset(g->mark, 0, (size_t)n);
  for (int v = 0; v < n; ++v) g->parent[v] = -1;
  queue_push(g, &tail, s);

  while (head < tail) {
    if (pops == RA_HUNG_MAX_ITER) {
      g->status |= RA_ST_BFS_CAP;
      if (g->stop_at_fatal) return -1;
    }
    int v = g->queue[head++];
    ++pops;
    g->mark[v] = 1;
    if (v == t) {
      found = 1;
      break;
    }
    for (int u = 0; u < n; ++u) {
      if (!g->mark[u] && g->residual[(size_t)v * n + u] > 0) {
        queue_push(g, &tail, u);
        g->parent[u] = v;
      }
    }
  }
  if (pops > g->max_pops) g->max_pops = pops;

  if (found) {
    float b = g->capacity[0];
    for (size_t k = 1; k < (size_t)n * n; ++k)
      if (g->capacity[k] > b) b = g->capacity[k];
    for (int v = t; g->parent[v] != -1; v = g->parent[v]) {
      float r = g->residual[(size_t)g->parent[v] * n + v];
      if (r < b) b = r;
    }
    for (int v = t; g->parent[v] != -1; v = g->parent[v]) {
      int p = g->parent[v];
      if (g->capacity[(size_t)p * n + v] > 0)
        g->flow[(size_t)p * n + v] += b;
      else
        g->flow[(size_t)v * n + p] -= b;
      g->residual[(size_t)p * n + v] -= b;
      g->residual[(size_t)v * n + p] += b;
    }
  }
  return found;
}

/* hungarian.cc:179-192 (MaxFlow) */
static int max_flow(flow_net *g) {
  const size_t nn = (size_t)g->n * g->n;
  memset(g->flow, 0, nn * sizeof(float));
  memcpy(g->residual, g->capacity, nn * sizeof(float));
  for (long i = 0;; ++i) {
    int r = augment(g);
    if (r < 0) return -1;
    if (!r) break;
    if (i == RA_HUNG_MAX_ITER) {
      g->status |= RA_ST_FLOW_CAP;
      if (g->stop_at_fatal) return -1;
    }
  }
  return 0;
}

/* hungarian.cc:194-217 (MaxBipartiteMatching): source -> X -> Y -> sink network
 * with unit capacities on the equality-graph edges; the X->Y flow is the matching. */
static int max_bipartite_matching(flow_net *g, const float *graph, int nx, int ny, float *matching) {
  const int n = g->n;
  const int s = 0, t = nx + ny + 1, x0 = 1, y0 = nx + 1;
  memset(g->capacity, 0, (size_t)n * n * sizeof(float));
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y) g->capacity[(size_t)(x0 + x) * n + (y0 + y)] = graph[x * ny + y];
  for (int x = 0; x < nx; ++x) g->capacity[(size_t)s * n + (x0 + x)] = 1.0f;
  for (int y = 0; y < ny; ++y) g->capacity[(size_t)(y0 + y) * n + t] = 1.0f;
  if (max_flow(g) < 0) return -1;
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y) matching[x * ny + y] = g->flow[(size_t)(x0 + x) * n + (y0 + y)];
  return 0;
}

/* hungarian.cc:219-248 (IsBipartiteMatchingSaturate) */
static int matching_saturates(const float *m, int nx, int ny) {
  if (nx >= ny) {
    for (int j = 0; j < ny; ++j) {
      float sum = 0;
      for (int i = 0; i < nx; ++i) sum += m[i * ny + j];
      if (sum == 0) return 0;
    }
  } else {
    for (int i = 0; i < nx; ++i) {
      float sum = 0;
      for (int j = 0; j < ny; ++j) sum += m[i * ny + j];
      if (sum == 0) return 0;
    }
  }
  return 1;
}

/* hungarian.cc:289-307 (GetMatchedX / GetMatchedY) */
static int matched_x_of(int y, const float *m, int nx, int ny) {
  for (int u = 0; u < nx; ++u)
    if (m[u * ny + y] == 1.0f) return u;
  return -1;
}
static int matched_y_of(int x, const float *m, int ny) {
  for (int v = 0; v < ny; ++v)
    if (m[x * ny + v] == 1.0f) return v;
  return -1;
}

/* hungarian.cc:309-325 (GetEqualityGraph) */
static void equality_graph(const float *w, const float *cx, const float *cy, int nx, int ny, float *eq) {
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y) {
      float d = cx[x] + cy[y] - w[x * ny + y];
      float ad = d > 0 ? d : -d;
      eq[x * ny + y] = ((double)ad <= RA_HUNG_EPS && (cx[x] > 0 || cy[y] > 0)) ? 1.0f : 0.0f;
    }
}

static int set_count(const unsigned char *s, int n) {
  int c = 0;
  for (int i = 0; i < n; ++i) c += s[i];
  return c;
}

/* hungarian.cc:335-488 (MinWeightedBipartiteCover) for one example. */
static int min_weighted_cover(const float *w, int nx, int ny, float *M, float *cx, float *cy, int stop_at_fatal,
                              long *max_pops, long *outer_rounds) {
  flow_net g;
  g.n = nx + ny + 2;
  const size_t nn = (size_t)g.n * g.n;
  g.capacity = (float *)malloc(nn * sizeof(float));
  g.flow = (float *)malloc(nn * sizeof(float));
  g.residual = (float *)malloc(nn * sizeof(float));
  g.mark = (unsigned char *)malloc((size_t)g.n);
  g.parent = (int *)malloc((size_t)g.n * sizeof(int));
  g.queue_cap = 4096;
  g.queue = (int *)malloc(g.queue_cap * sizeof(int));
  g.max_pops = 0;
  g.status = 0;
  g.stop_at_fatal = stop_at_fatal;

  float *eq = (float *)malloc((size_t)nx * ny * sizeof(float));
  unsigned char *S = (unsigned char *)calloc((size_t)nx, 1);
  unsigned char *T = (unsigned char *)calloc((size_t)ny, 1);
  unsigned char *NS = (unsigned char *)calloc((size_t)ny, 1);

  for (int x = 0; x < nx; ++x) {
    float mx = w[x * ny];
    for (int y = 1; y < ny; ++y)
      if (w[x * ny + y] > mx) mx = w[x * ny + y];
    cx[x] = mx;
  }
  for (int y = 0; y < ny; ++y) cy[y] = 0.0f;
  for (int k = 0; k < nx * ny; ++k) M[k] = 0.0f;

  int next_match = 1;
  int fatal = 0;
  long round = 0;
  for (;; ++round) {
    if (round == RA_HUNG_MAX_ITER) {
      g.status |= RA_ST_OUTER_CAP; /* reference: LOG(ERROR), return the unfinished matching */
      break;
    }
    equality_graph(w, cx, cy, nx, ny, eq);
    if (next_match) {
      if (max_bipartite_matching(&g, eq, nx, ny, M) < 0) {
        fatal = 1;
        break;
      }
      if (matching_saturates(M, nx, ny)) break;
      for (int u = 0; u < nx; ++u)
        if (matched_y_of(u, M, ny) == -1) {
          memset(S, 0, (size_t)nx);
          memset(T, 0, (size_t)ny);
          S[u] = 1;
          break;
        }
    }

    /* hungarian.cc:250-263 (GetSetBipartiteNeighbours) */
    memset(NS, 0, (size_t)ny);
    for (int x = 0; x < nx; ++x)
      if (S[x])
        for (int y = 0; y < ny; ++y)
          if (eq[x * ny + y] > 0) NS[y] = 1;

    if (memcmp(NS, T, (size_t)ny) == 0) {
      float a = FLT_MAX;
      for (int x = 0; x < nx; ++x)
        if (S[x])
          for (int y = 0; y < ny; ++y)
            if (!T[y]) {
              float d = cx[x] + cy[y] - w[x * ny + y];
              if (d < a) a = d;
            }
      if ((double)a < RA_HUNG_EPS) {
        next_match = 1;
        continue;
      }
      for (int x = 0; x < nx; ++x)
        if (S[x]) cx[x] -= a;
      for (int y = 0; y < ny; ++y)
        if (T[y]) cy[y] += a;
    } else {
      for (long j = 0; set_count(NS, ny) > set_count(T, ny); ++j) {
        if (j == RA_HUNG_MAX_ITER) {
          g.status |= RA_ST_EQUALIZE_CAP;
          if (stop_at_fatal) {
            fatal = 1;
            break;
          }
        }
        int y = -1;
        for (int v = 0; v < ny; ++v)
          if (NS[v] && !T[v]) {
            y = v;
            break;
          }
        int z = matched_x_of(y, M, nx, ny);
        if (z == -1) {
          next_match = 1;
          break;
        }
        next_match = 0;
        S[z] = 1;
        for (int v = 0; v < ny; ++v)
          if (eq[z * ny + v] > 0.0f) NS[v] = 1;
        T[y] = 1;
      }
      if (fatal) break;
    }
  }

  if (max_pops) *max_pops = g.max_pops;
  if (outer_rounds) *outer_rounds = round;
  int status = g.status;
  free(g.capacity);
  free(g.flow);
  free(g.residual);
  free(g.mark);
  free(g.parent);
  free(g.queue);
  free(eq);
  free(S);
  free(T);
  free(NS);
  return status;
}

/*
 * Batch entry point, hungarian.cc:506-537 (ComputeHungarianBatch); B == 1 covers
 * the rank-2 form (hungarian.cc:490-504).  Row-major fp32:
 *   W [B,nx,ny] -> M [B,nx,ny], cx [B,nx], cy [B,ny]; status[B] gets the RA_ST_* bits,
 *   max_pops[B] (optional) the largest BFS pop count per example.
 * Returns the OR of all status words.
 */
int ra_oracle_hungarian_f32(const float *W, int B, int nx, int ny, float *M, float *cx, float *cy, int *status,
                            long *max_pops, int stop_at_fatal) {
  int all = 0;
  for (int b = 0; b < B; ++b) {
    long pops = 0;
    int st = min_weighted_cover(W + (size_t)b * nx * ny, nx, ny, M + (size_t)b * nx * ny, cx + (size_t)b * nx,
                                cy + (size_t)b * ny, stop_at_fatal, &pops, NULL);
    if (status) status[b] = st;
    if (max_pops) max_pops[b] = pops;
    all |= st;
  }
  return all;
}
