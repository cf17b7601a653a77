/*
 * mps_gnat.c — CPU restatement of the nearest-neighbour structure the reference uses for its tree:
 * ompl::NearestNeighborsGNAT<MotionPtr> (Geometric Near-neighbour Access Tree, S. Brin 1995), created with
 * OMPL's default parameters at src/mps/planner/pushing/algorithm/RRT.cpp:46-51 (degree 8, min 4, max 12,
 * 50 points per leaf) and queried with nearest() at :202.  OMPL is un-vendored and un-pinned in the reference,
 * so this follows its published algorithm: k-centres pivots on leaf overflow, per-child [min, max] range
 * tables, best-first search with range pruning.  Exact: it returns a true nearest neighbour (ties: any).
 *
 * TEST / BENCH INFRASTRUCTURE ONLY (CPU baseline of the NN queries); never linked into the product.
 * Differences from OMPL that do not affect exactness: the first k-centres pivot is the first point instead of
 * a random one, children are visited in index order instead of a rotating permutation, no removal support.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mps_oracle.h"

#define GNAT_DEGREE 8
#define GNAT_MIN_DEGREE 4
#define GNAT_MAX_DEGREE 12
#define GNAT_MAX_LEAF 50

typedef struct gnat_node {
    int pivot;          /* point id */
    int degree;         /* target number of children when this leaf splits */
    double min_radius, max_radius;
    double* min_range;  /* [n_children] distances from this node's pivot to the points of child j */
    double* max_range;
    int* data;          /* leaf points */
    int n_data, cap_data;
    struct gnat_node** children;
    int n_children;
} gnat_node;

struct orc_gnat {
    const mps_scene* sc;
    float weights[MPS_MAX_BODIES];
    int has_weights;
    uint32_t mask;
    float* pts; /* [n][3B] */
    int n, cap;
    gnat_node* root;
    long n_dist; /* distance evaluations (for reporting) */
};

static double gdist(orc_gnat* g, const float* a, const float* b)
{
    g->n_dist++;
    return orc_distance(g->sc, a, b, g->has_weights ? g->weights : 0, g->mask);
}

static const float* gpt(const orc_gnat* g, int id) { return g->pts + (size_t)id * 3 * (size_t)g->sc->n_bodies; }

static gnat_node* node_new(int degree, int cap_range, int pivot)
{
    gnat_node* n = (gnat_node*)calloc(1, sizeof(gnat_node));
    n->pivot = pivot;
    n->degree = degree;
    n->min_radius = INFINITY;
    n->max_radius = -INFINITY;
    n->min_range = (double*)malloc(sizeof(double) * (size_t)cap_range);
    n->max_range = (double*)malloc(sizeof(double) * (size_t)cap_range);
    for (int i = 0; i < cap_range; ++i) {
        n->min_range[i] = INFINITY;
        n->max_range[i] = -INFINITY;
    }
    return n;
}

static void node_free(gnat_node* n)
{
    if (!n) return;
    for (int i = 0; i < n->n_children; ++i) node_free(n->children[i]);
    free(n->children);
    free(n->min_range);
    free(n->max_range);
    free(n->data);
    free(n);
}

static void update_radius(gnat_node* n, double d)
{
    if (n->min_radius > d) n->min_radius = d;
    if (n->max_radius < d) n->max_radius = d;
}

static void update_range(gnat_node* n, int i, double d)
{
    if (n->min_range[i] > d) n->min_range[i] = d;
    if (n->max_range[i] < d) n->max_range[i] = d;
}

static void leaf_push(gnat_node* n, int id)
{
    if (n->n_data == n->cap_data) {
        n->cap_data = n->cap_data ? 2 * n->cap_data : 16;
        n->data = (int*)realloc(n->data, sizeof(int) * (size_t)n->cap_data);
    }
    n->data[n->n_data++] = id;
}

static void node_add(orc_gnat* g, gnat_node* n, int id);

/* leaf overflow: choose `degree` pivots by greedy k-centres, build the children, distribute the points */
static void node_split(orc_gnat* g, gnat_node* n)
{
    int m = n->n_data, k = n->degree, i, j, c;
    if (k > m) k = m;
    const int ks = k; /* row stride of dists */
    int* centers = (int*)malloc(sizeof(int) * (size_t)k);
    double* dists = (double*)malloc(sizeof(double) * (size_t)m * (size_t)k); /* dists[i * k + c] */
    double* mind = (double*)malloc(sizeof(double) * (size_t)m);
    for (i = 0; i < m; ++i) mind[i] = INFINITY;
    centers[0] = 0;
    int nc = 1;
    for (c = 1; c <= k; ++c) {
        int last = centers[c - 1], far = 0;
        double best = -1.0;
        for (i = 0; i < m; ++i) {
            double d = gdist(g, gpt(g, n->data[i]), gpt(g, n->data[last]));
            dists[(size_t)i * ks + (c - 1)] = d;
            if (d < mind[i]) mind[i] = d;
            if (mind[i] > best) {
                best = mind[i];
                far = i;
            }
        }
        if (c == k || best <= 0.0) break; /* all remaining points coincide with a centre */
        centers[c] = far;
        nc = c + 1;
    }
    k = nc;
    n->children = (gnat_node**)calloc((size_t)k, sizeof(gnat_node*));
    n->n_children = k;
    for (c = 0; c < k; ++c) {
        gnat_node* ch = node_new(GNAT_DEGREE, k, n->data[centers[c]]);
        n->children[c] = ch;
    }
    n->degree = k;
    for (i = 0; i < m; ++i) {
        int is_center = 0;
        for (c = 0; c < k; ++c)
            if (centers[c] == i) is_center = 1;
        /* closest centre */
        int best = 0;
        for (c = 1; c < k; ++c)
            if (dists[(size_t)i * ks + c] < dists[(size_t)i * ks + best]) best = c;
        gnat_node* ch = n->children[best];
        if (!is_center) {
            leaf_push(ch, n->data[i]);
            update_radius(ch, dists[(size_t)i * ks + best]);
        }
        for (j = 0; j < k; ++j) update_range(n->children[j], best, dists[(size_t)i * ks + j]);
    }
    /* OMPL adapts the children's degree to their share of the points */
    for (c = 0; c < k; ++c) {
        gnat_node* ch = n->children[c];
        int deg = (int)ceil((double)n->degree * (double)ch->n_data / (double)(m > 0 ? m : 1));
        if (deg < GNAT_MIN_DEGREE) deg = GNAT_MIN_DEGREE;
        if (deg > GNAT_MAX_DEGREE) deg = GNAT_MAX_DEGREE;
        ch->degree = deg;
    }
    free(n->data);
    n->data = 0;
    n->n_data = n->cap_data = 0;
    free(centers);
    free(dists);
    free(mind);
    /* children that overflowed split in turn */
    for (c = 0; c < k; ++c) {
        gnat_node* ch = n->children[c];
        if (ch->n_data > GNAT_MAX_LEAF && ch->n_data > ch->degree) node_split(g, ch);
    }
}

static void node_add(orc_gnat* g, gnat_node* n, int id)
{
    while (n->n_children > 0) {
        int k = n->n_children, best = 0, c;
        double d[GNAT_MAX_DEGREE];
        for (c = 0; c < k; ++c) d[c] = gdist(g, gpt(g, id), gpt(g, n->children[c]->pivot));
        for (c = 1; c < k; ++c)
            if (d[c] < d[best]) best = c;
        for (c = 0; c < k; ++c) update_range(n->children[c], best, d[c]);
        update_radius(n->children[best], d[best]);
        n = n->children[best];
    }
    leaf_push(n, id);
    if (n->n_data > GNAT_MAX_LEAF && n->n_data > n->degree) node_split(g, n);
}

orc_gnat* orc_gnat_create(const mps_scene* sc, const float* weights, uint32_t mask)
{
    orc_gnat* g = (orc_gnat*)calloc(1, sizeof(orc_gnat));
    g->sc = sc;
    g->mask = mask;
    if (weights) {
        memcpy(g->weights, weights, sizeof(float) * (size_t)sc->n_bodies);
        g->has_weights = 1;
    }
    return g;
}

void orc_gnat_free(orc_gnat* g)
{
    if (!g) return;
    node_free(g->root);
    free(g->pts);
    free(g);
}

int orc_gnat_add(orc_gnat* g, const float* state)
{
    size_t dim = 3 * (size_t)g->sc->n_bodies;
    if (g->n == g->cap) {
        g->cap = g->cap ? 2 * g->cap : 1024;
        g->pts = (float*)realloc(g->pts, sizeof(float) * dim * (size_t)g->cap);
    }
    memcpy(g->pts + dim * (size_t)g->n, state, sizeof(float) * dim);
    int id = g->n++;
    if (!g->root) {
        g->root = node_new(GNAT_DEGREE, GNAT_MAX_DEGREE, id);
    } else {
        node_add(g, g->root, id);
    }
    return id;
}

int orc_gnat_size(const orc_gnat* g) { return g->n; }
long orc_gnat_distance_evaluations(const orc_gnat* g) { return g->n_dist; }

/* best-first search; a binary min-heap of (node, distance to its pivot) keyed by dist - max_radius */
typedef struct heap_item {
    gnat_node* node;
    double dist_to_pivot, key;
} heap_item;

typedef struct heap {
    heap_item* a;
    int n, cap;
} heap;

static void heap_push(heap* h, heap_item it)
{
    if (h->n == h->cap) {
        h->cap = h->cap ? 2 * h->cap : 64;
        h->a = (heap_item*)realloc(h->a, sizeof(heap_item) * (size_t)h->cap);
    }
    int i = h->n++;
    h->a[i] = it;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (h->a[p].key <= h->a[i].key) break;
        heap_item t = h->a[p];
        h->a[p] = h->a[i];
        h->a[i] = t;
        i = p;
    }
}

static heap_item heap_pop(heap* h)
{
    heap_item top = h->a[0];
    h->a[0] = h->a[--h->n];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < h->n && h->a[l].key < h->a[m].key) m = l;
        if (r < h->n && h->a[r].key < h->a[m].key) m = r;
        if (m == i) break;
        heap_item t = h->a[m];
        h->a[m] = h->a[i];
        h->a[i] = t;
        i = m;
    }
    return top;
}

int orc_gnat_nearest(orc_gnat* g, const float* query, double* dist_out)
{
    if (!g->root) {
        if (dist_out) *dist_out = INFINITY;
        return -1;
    }
    heap h = {0, 0, 0};
    int best = g->root->pivot;
    double bd = gdist(g, query, gpt(g, best));
    heap_item it = {g->root, bd, bd - g->root->max_radius};
    heap_push(&h, it);
    while (h.n > 0) {
        heap_item cur = heap_pop(&h);
        gnat_node* n = cur.node;
        if (n != g->root &&
            (cur.dist_to_pivot > n->max_radius + bd || cur.dist_to_pivot < n->min_radius - bd))
            continue;
        for (int i = 0; i < n->n_data; ++i) {
            double d = gdist(g, query, gpt(g, n->data[i]));
            if (d < bd || (d == bd && n->data[i] < best)) {
                bd = d;
                best = n->data[i];
            }
        }
        if (n->n_children > 0) {
            int k = n->n_children, alive[GNAT_MAX_DEGREE], i, j;
            double d[GNAT_MAX_DEGREE];
            for (i = 0; i < k; ++i) alive[i] = 1;
            for (i = 0; i < k; ++i) {
                if (!alive[i]) continue;
                gnat_node* ch = n->children[i];
                d[i] = gdist(g, query, gpt(g, ch->pivot));
                if (d[i] < bd || (d[i] == bd && ch->pivot < best)) {
                    bd = d[i];
                    best = ch->pivot;
                }
                for (j = 0; j < k; ++j)
                    if (alive[j] && i != j && (d[i] - bd > ch->max_range[j] || d[i] + bd < ch->min_range[j])) alive[j] = 0;
            }
            for (i = 0; i < k; ++i) {
                if (!alive[i]) continue;
                gnat_node* ch = n->children[i];
                if (d[i] - bd <= ch->max_radius && d[i] + bd >= ch->min_radius) {
                    heap_item c = {ch, d[i], d[i] - ch->max_radius};
                    heap_push(&h, c);
                }
            }
        }
    }
    free(h.a);
    if (dist_out) *dist_out = bd;
    return best;
}
