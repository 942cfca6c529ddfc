/*
 * hostgen.c -- host-side synthetic catchment generator (input generation for tests and bench).
 *
 * SPHY never creates an LDD: it reads one (/root/reference/modules/routing.py:34).  PCRaster's
 * lddcreate is not available here, so synthetic catchments get their local drain direction map
 * from this priority-flood: cells are flooded from the outlet seeds in order of (filled elevation,
 * index); every cell drains to its steepest strictly-lower neighbour on the filled surface
 * (ties -> lowest LDD code) and, inside filled depressions / flats, to the cell that flooded it.
 * Strictly-descending edges cannot close a cycle and flood-parent edges form a forest, so the
 * result is a sound LDD (acyclic, every cell reaches a pit).
 *
 * Build: gcc -O2 -fPIC -shared -o sphy_b200/_hostgen.so sphy_b200/csrc/hostgen.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float z; int32_t i; } hnode;

static inline int hless(hnode a, hnode b) { return a.z < b.z || (a.z == b.z && a.i < b.i); }

static void hpush(hnode *h, int64_t *n, hnode v)
{
    int64_t k = (*n)++;
    while (k > 0) {
        int64_t p = (k - 1) >> 1;
        if (!hless(v, h[p])) break;
        h[k] = h[p];
        k = p;
    }
    h[k] = v;
}

static hnode hpop(hnode *h, int64_t *n)
{
    hnode top = h[0], v = h[--(*n)];
    int64_t k = 0, m = *n;
    for (;;) {
        int64_t c = 2 * k + 1;
        if (c >= m) break;
        if (c + 1 < m && hless(h[c + 1], h[c])) c++;
        if (!hless(h[c], v)) break;
        h[k] = h[c];
        k = c;
    }
    if (m > 0) h[k] = v;
    return top;
}

/* LDD keypad code for a (drow, dcol) step */
static inline uint8_t code_of(int dr, int dc) { return (uint8_t)(5 - 3 * dr + dc); }

/*
 * dem[nrows*ncols] float32 (modified in place to the filled surface).
 * seed_mode 0: every border cell is an outlet seed; 1: single outlet at (seed_r, seed_c).
 * Outputs: ldd (uint8 codes), slope (drop/distance along the chosen direction, clipped to [1e-4,1]).
 * Returns 0, or -1 on allocation failure.
 */
int hostgen_flood_ldd(float *dem, int nrows, int ncols, float cellsize, int seed_mode, int seed_r,
                      int seed_c, uint8_t *ldd, float *slope)
{
    static const int NR[8] = {1, 1, 1, 0, 0, -1, -1, -1};
    static const int NC[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
    int64_t ncell = (int64_t)nrows * ncols;
    hnode *heap = (hnode *)malloc(sizeof(hnode) * (size_t)ncell);
    int32_t *parent = (int32_t *)malloc(sizeof(int32_t) * (size_t)ncell);
    uint8_t *closed = (uint8_t *)calloc((size_t)ncell, 1);
    if (!heap || !parent || !closed) { free(heap); free(parent); free(closed); return -1; }
    int64_t hn = 0;
    if (seed_mode == 1) {
        int64_t g = (int64_t)seed_r * ncols + seed_c;
        closed[g] = 1; parent[g] = -1;
        hpush(heap, &hn, (hnode){dem[g], (int32_t)g});
    } else {
        for (int r = 0; r < nrows; ++r)
            for (int c = 0; c < ncols; ++c)
                if (r == 0 || c == 0 || r == nrows - 1 || c == ncols - 1) {
                    int64_t g = (int64_t)r * ncols + c;
                    closed[g] = 1; parent[g] = -1;
                    hpush(heap, &hn, (hnode){dem[g], (int32_t)g});
                }
    }
    while (hn > 0) {
        hnode t = hpop(heap, &hn);
        int r = t.i / ncols, c = t.i % ncols;
        for (int k = 0; k < 8; ++k) {
            int rr = r + NR[k], cc = c + NC[k];
            if (rr < 0 || rr >= nrows || cc < 0 || cc >= ncols) continue;
            int64_t g = (int64_t)rr * ncols + cc;
            if (closed[g]) continue;
            closed[g] = 1;
            parent[g] = t.i;
            if (dem[g] < t.z) dem[g] = t.z;          /* fill */
            hpush(heap, &hn, (hnode){dem[g], (int32_t)g});
        }
    }
    for (int r = 0; r < nrows; ++r)
        for (int c = 0; c < ncols; ++c) {
            int64_t g = (int64_t)r * ncols + c;
            if (parent[g] < 0) { ldd[g] = 5; slope[g] = 1e-4f; continue; }
            float best = 0.0f; int bk = -1; uint8_t bcode = 10;
            for (int k = 0; k < 8; ++k) {
                int rr = r + NR[k], cc = c + NC[k];
                if (rr < 0 || rr >= nrows || cc < 0 || cc >= ncols) continue;
                float drop = dem[g] - dem[(int64_t)rr * ncols + cc];
                if (drop <= 0.0f) continue;
                float dist = (NR[k] != 0 && NC[k] != 0) ? 1.41421356f * cellsize : cellsize;
                float s = drop / dist;
                uint8_t code = code_of(NR[k], NC[k]);
                if (s > best || (s == best && code < bcode)) { best = s; bk = k; bcode = code; }
            }
            if (bk >= 0) {
                ldd[g] = bcode;
            } else {
                int pr = parent[g] / ncols, pc = parent[g] % ncols;
                ldd[g] = code_of(pr - r, pc - c);
                best = 0.0f;
            }
            slope[g] = best < 1e-4f ? 1e-4f : (best > 1.0f ? 1.0f : best);
        }
    free(heap); free(parent); free(closed);
    return 0;
}

/* splitmix64 -> uniform float in [0,1) */
static inline uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline float u01(uint64_t seed, int64_t a, int64_t b)
{
    return (float)(mix64(seed ^ mix64((uint64_t)a * 0x100000001B3ull + (uint64_t)b)) >> 40) * (1.0f / 16777216.0f);
}
/* smooth value noise on a lattice of the given wavelength (cells) */
static float vnoise(uint64_t seed, float r, float c, float wl)
{
    float fr = r / wl, fc = c / wl;
    int64_t r0 = (int64_t)floorf(fr), c0 = (int64_t)floorf(fc);
    float tr = fr - (float)r0, tc = fc - (float)c0;
    tr = tr * tr * (3.0f - 2.0f * tr);
    tc = tc * tc * (3.0f - 2.0f * tc);
    float a = u01(seed, r0, c0), b = u01(seed, r0, c0 + 1), d = u01(seed, r0 + 1, c0), e = u01(seed, r0 + 1, c0 + 1);
    return (a * (1 - tc) + b * tc) * (1 - tr) + (d * (1 - tc) + e * tc) * tr - 0.5f;
}

/*
 * Synthetic DEM: a 2000 m tilt that converges on the outlet at the middle of the
 * south edge, three octaves of value noise (300/100/30 m at N/4, N/16, N/64) and U(0,1) m jitter.
 * zmin/zmax rescale the result (config 2 uses 1500..6500 m); pass zmin>=zmax to skip.
 */
void hostgen_dem(float *dem, int nrows, int ncols, uint64_t seed, float zmin, float zmax)
{
    float N = (float)(nrows > ncols ? nrows : ncols);
    float orow = (float)(nrows - 1), ocol = 0.5f * (float)(ncols - 1);
    float R = sqrtf(orow * orow + ocol * ocol);
    float lo = 1e30f, hi = -1e30f;
    for (int r = 0; r < nrows; ++r)
        for (int c = 0; c < ncols; ++c) {
            float dr = (float)r - orow, dc = (float)c - ocol;
            float dist = sqrtf(dr * dr + dc * dc);
            float z = 2000.0f * (dist / R);
            z += 600.0f * vnoise(seed + 1, (float)r, (float)c, N / 4.0f);
            z += 200.0f * vnoise(seed + 2, (float)r, (float)c, N / 16.0f);
            z += 60.0f * vnoise(seed + 3, (float)r, (float)c, N / 64.0f > 1.0f ? N / 64.0f : 1.0f);
            z += u01(seed + 4, r, c);
            dem[(int64_t)r * ncols + c] = z;
            if (z < lo) lo = z;
            if (z > hi) hi = z;
        }
    if (zmax > zmin && hi > lo) {
        float s = (zmax - zmin) / (hi - lo);
        int64_t ncell = (int64_t)nrows * ncols;
        for (int64_t i = 0; i < ncell; ++i) dem[i] = zmin + (dem[i] - lo) * s;
    }
}

/* Upstream cell counts (incl. self) of a full-raster LDD; used only to place synthetic stations,
 * reservoirs and lakes.  Returns 0, -1 on a cycle. */
int hostgen_nup(const uint8_t *ldd, int nrows, int ncols, int32_t *nup)
{
    static const int DRr[10] = {0, 1, 1, 1, 0, 0, 0, -1, -1, -1};
    static const int DCc[10] = {0, -1, 0, 1, -1, 0, 1, -1, 0, 1};
    int64_t ncell = (int64_t)nrows * ncols;
    int32_t *down = (int32_t *)malloc(sizeof(int32_t) * (size_t)ncell);
    int32_t *queue = (int32_t *)malloc(sizeof(int32_t) * (size_t)ncell);
    uint8_t *deg = (uint8_t *)calloc((size_t)ncell, 1);
    if (!down || !queue || !deg) { free(down); free(queue); free(deg); return -2; }
    for (int r = 0; r < nrows; ++r)
        for (int c = 0; c < ncols; ++c) {
            int64_t g = (int64_t)r * ncols + c;
            int code = ldd[g];
            down[g] = -1;
            nup[g] = 1;
            if (code >= 1 && code <= 9 && code != 5) {
                int rr = r + DRr[code], cc = c + DCc[code];
                if (rr >= 0 && rr < nrows && cc >= 0 && cc < ncols) down[g] = (int32_t)((int64_t)rr * ncols + cc);
            }
        }
    for (int64_t g = 0; g < ncell; ++g) if (down[g] >= 0) deg[down[g]]++;
    int64_t head = 0, tail = 0;
    for (int64_t g = 0; g < ncell; ++g) if (!deg[g]) queue[tail++] = (int32_t)g;
    while (head < tail) {
        int32_t g = queue[head++], d = down[g];
        if (d >= 0) { nup[d] += nup[g]; if (--deg[d] == 0) queue[tail++] = d; }
    }
    free(down); free(queue); free(deg);
    return tail == ncell ? 0 : -1;
}
