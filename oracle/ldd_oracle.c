/*
 * oracle/ldd_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never on the product path).
 *
 * Plain-C restatement of the PCRaster LDD operations SPHY's routing calls:
 *   pcr.accuflux(ldd, m)                /root/reference/modules/routing.py:27
 *   pcr.accufractionflux(ldd, m, f)     /root/reference/modules/advanced_routing.py:29
 *   pcr.upstream(ldd, x)                /root/reference/modules/advanced_routing.py:35,159
 *   pcr.catchmenttotal(x, ldd)          /root/reference/sphy.py:838 (== accuflux with args swapped)
 * and of the canonical integer structures of SURVEY.md Appendix C
 * (down, indeg, level, order, level_ptr, donor CSR, nup).
 *
 * PCRaster itself is a third-party dependency that is NOT in /root/reference and is
 * not installable here (version unpinned by the reference) -> "parity unpinned" at the
 * PCRaster boundary.  Semantics restated from PCRaster's published documentation:
 *   - LDD codes are the numeric keypad (7 8 9 / 4 5 6 / 1 2 3), 5 = pit, row index grows south.
 *   - accuflux: out-flux of a cell = its material + the out-flux of every cell draining into it.
 *   - Each cell's sum (own material + <=8 donor fluxes) is formed in a REAL8 local and
 *     stored as REAL4 (assumption A1 in DESIGN.md): flux = (float)((double)m + sum (double)flux_donor).
 *   - accufractionflux: flux = fraction * (material + inflow); state = the remainder.
 *   - upstream: sum of x over the direct donors (0 if none).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o oracle/_build/libldd_oracle.so oracle/ldd_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const int DR[10] = {0, 1, 1, 1, 0, 0, 0, -1, -1, -1};
static const int DC[10] = {0, -1, 0, 1, -1, 0, 1, -1, 0, 1};

/* Number of active cells and raster->compact index map (row-major order).
 * A cell is active when its LDD code is 1..9 and (active==NULL or active[i]!=0). */
int64_t ldd_compact(const uint8_t *ldd, const uint8_t *active, int nrows, int ncols, int32_t *cidx)
{
    int64_t n = 0, ncell = (int64_t)nrows * ncols;
    for (int64_t i = 0; i < ncell; ++i) {
        int ok = ldd[i] >= 1 && ldd[i] <= 9 && (!active || active[i]);
        cidx[i] = ok ? (int32_t)n++ : -1;
    }
    return n;
}

/* Build every Appendix-C structure.  Arrays are caller-allocated:
 *   down[n] indeg[n] level[n] order[n] level_ptr[n+2] don_ptr[n+1] don_idx[n] nup[n]
 * Returns the number of levels (>=1), or -1 on a cycle, -2 on allocation failure. */
int32_t ldd_build(const uint8_t *ldd, const int32_t *cidx, int nrows, int ncols, int64_t n,
                  int32_t *down, uint8_t *indeg, int32_t *level, int32_t *order,
                  int32_t *level_ptr, int32_t *don_ptr, int32_t *don_idx, int64_t *nup)
{
    memset(indeg, 0, (size_t)n);
    for (int r = 0; r < nrows; ++r)
        for (int c = 0; c < ncols; ++c) {
            int64_t g = (int64_t)r * ncols + c;
            int32_t i = cidx[g];
            if (i < 0) continue;
            int code = ldd[g];
            int32_t d = -1;
            if (code != 5) {
                int rr = r + DR[code], cc = c + DC[code];
                if (rr >= 0 && rr < nrows && cc >= 0 && cc < ncols)
                    d = cidx[(int64_t)rr * ncols + cc];
            }
            down[i] = d;
        }
    for (int64_t i = 0; i < n; ++i)
        if (down[i] >= 0) indeg[down[i]]++;
    /* donor CSR, donors sorted by ascending compact index (scan i ascending) */
    don_ptr[0] = 0;
    for (int64_t i = 0; i < n; ++i) don_ptr[i + 1] = don_ptr[i] + indeg[i];
    int32_t *fill = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    if (!fill) return -2;
    for (int64_t i = 0; i < n; ++i) fill[i] = don_ptr[i];
    for (int64_t i = 0; i < n; ++i)
        if (down[i] >= 0) don_idx[fill[down[i]]++] = (int32_t)i;
    /* Kahn peeling: level = longest upstream path */
    uint8_t *rem = (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
    if (!rem) { free(fill); return -2; }
    memcpy(rem, indeg, (size_t)n);
    int64_t head = 0, tail = 0;
    for (int64_t i = 0; i < n; ++i) {
        level[i] = 0;
        nup[i] = 1;
        if (indeg[i] == 0) order[tail++] = (int32_t)i;
    }
    int32_t nlev = 0;
    level_ptr[0] = 0;
    while (head < tail) {
        int64_t lev_end = tail;           /* [head, lev_end) is one level, ascending index */
        level_ptr[++nlev] = (int32_t)lev_end;
        /* receivers completed by this level join the next one; collect then sort by index
         * (they are discovered in donor order, not index order) */
        int64_t next_begin = tail;
        for (int64_t k = head; k < lev_end; ++k) {
            int32_t i = order[k], d = down[i];
            if (d >= 0) {
                nup[d] += nup[i];
                if (--rem[d] == 0) { level[d] = nlev; order[tail++] = d; }
            }
        }
        /* ascending compact index inside the level: simple LSD radix is overkill; use qsort */
        if (tail - next_begin > 1) {
            int32_t *seg = order + next_begin;
            int64_t m = tail - next_begin;
            /* insertion into a bitmap would be O(n) per level; comb sort keeps this dependency-free */
            int64_t gap = m;
            int swapped = 1;
            while (gap > 1 || swapped) {
                gap = (gap * 10) / 13;
                if (gap < 1) gap = 1;
                swapped = 0;
                for (int64_t a = 0; a + gap < m; ++a)
                    if (seg[a] > seg[a + gap]) {
                        int32_t t = seg[a]; seg[a] = seg[a + gap]; seg[a + gap] = t; swapped = 1;
                    }
            }
        }
        head = lev_end;
    }
    free(fill);
    free(rem);
    if (tail != n) return -1;             /* cycle in the LDD */
    level_ptr[nlev + 1] = (int32_t)n;     /* sentinel, SURVEY Appendix C: level_ptr[L+2] */
    return nlev;
}

/* accuflux / accufractionflux (frac may be NULL == 1 everywhere).
 * flux[i] = (float)(frac[i] * ((double)m[i] + sum_donors (double)flux[d])); state optional. */
void ldd_accuflux(const int32_t *order, const int32_t *down, int64_t n, const float *m,
                  const float *frac, float *flux, float *state)
{
    double *acc = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) acc[i] = (double)m[i];
    for (int64_t k = 0; k < n; ++k) {
        int32_t i = order[k];
        double tot = acc[i];
        float f;
        if (frac) {
            f = (float)((double)frac[i] * tot);
            if (state) state[i] = (float)(tot - (double)f);
        } else {
            f = (float)tot;
            if (state) state[i] = 0.0f;
        }
        flux[i] = f;
        if (down[i] >= 0) acc[down[i]] += (double)f;
    }
    free(acc);
}

/* accucapacityflux / accucapacitystate [PCR-doc]: what arrives (own material + donor fluxes, formed in double like
 * accuflux) leaves up to the cell's transport capacity; the rest stays behind as the state. */
void ldd_accucapacity(const int32_t *order, const int32_t *down, int64_t n, const float *m, const float *cap,
                      float *flux, float *state)
{
    double *acc = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) acc[i] = (double)m[i];
    for (int64_t k = 0; k < n; ++k) {
        int32_t i = order[k];
        double tot = acc[i];
        float f = tot < (double)cap[i] ? (float)tot : cap[i];
        if (!(tot == tot) || !(cap[i] == cap[i])) f = (float)(tot + (double)cap[i]);   /* MV in -> MV out */
        flux[i] = f;
        if (state) state[i] = (float)(tot - (double)f);
        if (down[i] >= 0) acc[down[i]] += (double)f;
    }
    free(acc);
}

/* The other candidate for PCRaster's accumulation arithmetic (assumption A1 as first written): the inflow of a
 * cell is a REAL4 variable to which the donors' fluxes are added one by one, in traversal order (here: routing order =
 * level, then cell index), each add rounded to REAL4.  Identical to ldd_accuflux wherever a cell has at most one donor;
 * at confluences it rounds once per donor instead of once per cell.  Unverifiable without a PCRaster build; kept so
 * that every routing path can report its distance to both candidate reference models. */
void ldd_accuflux_f32inc(const int32_t *order, const int32_t *down, int64_t n, const float *m, float *flux)
{
    for (int64_t i = 0; i < n; ++i) flux[i] = m[i];
    for (int64_t k = 0; k < n; ++k) {
        int32_t i = order[k];
        if (down[i] >= 0) {
            volatile float s = flux[down[i]] + flux[i];      /* one REAL4 add per donor */
            flux[down[i]] = s;
        }
    }
}

/* float64 twin used for error attribution (no REAL4 rounding along the path) */
void ldd_accuflux_f64(const int32_t *order, const int32_t *down, int64_t n, const float *m,
                      const float *frac, double *flux)
{
    for (int64_t i = 0; i < n; ++i) flux[i] = (double)m[i];
    for (int64_t k = 0; k < n; ++k) {
        int32_t i = order[k];
        if (frac) flux[i] *= (double)frac[i];
        if (down[i] >= 0) flux[down[i]] += flux[i];
    }
}

/* upstream(ldd, x): sum over direct donors, formed in double, stored REAL4 */
void ldd_upstream(const int32_t *don_ptr, const int32_t *don_idx, int64_t n, const float *x, float *out)
{
    for (int64_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int32_t k = don_ptr[i]; k < don_ptr[i + 1]; ++k) s += (double)x[don_idx[k]];
        out[i] = (float)s;
    }
}
