/*
 * kmc_oracle.c -- CPU restatement of the kmc_dn hop loop.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the B200 hop-loop kernels.  It is never
 * linked into, imported by, or called from the product (kmc_dn_b200/).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.
 *
 * Two reference semantics are restated, each from the reference's own sources
 * (paths relative to /root/reference):
 *
 *   (A) "py"  -- the numba loop  kmc_dopant_networks.py:33-135
 *                (_simulate_discrete_record) and :137-163 (_transition_possible).
 *                fp64 throughout, energies recomputed from scratch every hop,
 *                S*S cumulative list incl. zero entries, normalised compare.
 *                PINNED: checked bit-for-bit against the unmodified numba
 *                function run in the build container (oracle/make_golden.py,
 *                tests/golden/py_replay_*.npz).
 *
 *   (B) "go"  -- the Go loop  goSimulation/simulation.go:194-325 (simulate),
 *                :327-432 (simulateRecordPlus), :40-55 (transition_possible),
 *                :58-80 (calcTransitionList), :107-130 (makeJump),
 *                :163-188 (getRandomEvent), :29-38 (getKey) and the narrowing
 *                wrappers goSimulation/simulationWrapper.go:37-56,83-169,250-316.
 *                fp32 rates/energies/cumulative list, fp64 time.
 *                Go toolchain is absent from the build image, so the hop
 *                sequence of this mode is "parity unpinned" by an executable
 *                reference; it is pinned statistically by the reference's 400
 *                .kmc fixtures (generated with wrapperSimulateRecordPlus,
 *                thesis_indrek/generate_tests.py:45-51).
 *
 * Random numbers are INJECTED: the caller passes the per-hop variates so that
 * any implementation can be replayed against the same stream.
 *   py: u[2k]   -> dwell time  -log(1-u)*(1/total)   (kmc_dopant_networks.py:100)
 *       u[2k+1] -> event pick                         (kmc_dopant_networks.py:106)
 *   go: e[k]    -> unit exponential variate (rand.ExpFloat64, simulation.go:297)
 *       u[k]    -> float32 uniform in [0,1) (rand.Float32,    simulation.go:164)
 * When the stream pointers are NULL an internal xoshiro256++ generator is used
 * (statistical runs and CPU-baseline timing).
 *
 * Build:  gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared -pthread
 *         (-ffp-contract=off: amd64 Go never fuses x*y+z)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* ------------------------------------------------------------------ RNG -- */
typedef struct { uint64_t s[4]; } xo_t;
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static void xo_seed(xo_t *g, uint64_t seed) {
    for (int i = 0; i < 4; i++) g->s[i] = splitmix64(&seed);
}
static inline uint64_t xo_next(xo_t *g) {
    uint64_t *s = g->s;
    uint64_t r = rotl64(s[0] + s[3], 23) + s[0];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t; s[3] = rotl64(s[3], 45);
    return r;
}
static inline double xo_double(xo_t *g) { return (double)(xo_next(g) >> 11) * (1.0 / 9007199254740992.0); }
static inline float xo_float(xo_t *g) { return (float)(xo_next(g) >> 40) * (1.0f / 16777216.0f); }
static inline double xo_exp(xo_t *g) { return -log(1.0 - xo_double(g)); }

/* --------------------------------------------------- shared predicates -- */
/* simulation.go:40-55 / kmc_dopant_networks.py:137-163 (same truth table). */
static inline int transition_possible(int i, int j, int N, const uint8_t *occ) {
    if (i == j) return 0;
    if (i >= N && j >= N) return 0;
    if (i >= N) return !occ[j];
    if (j >= N) return occ[i];
    return occ[i] && !occ[j];
}

/* ===================================================================== (A) */
/*
 * kmc_oracle_py -- restatement of _simulate_discrete_record
 * (kmc_dopant_networks.py:33-135).  All arrays fp64, row-major.
 *   occupation      uint8[N]   in/out   (honoured, like the numba function)
 *   site_energies   f64[S]     in: [N:] electrode energies; out: [:N] last energies
 *   electrode_occ   int64[P]   in/out   (accumulated, not zeroed: :119,:123)
 *   u               f64[2*hops] injected uniforms, or NULL (internal RNG, seed)
 *   traffic         f64[S*S]   out or NULL (zeroed here, :53) -- only if record
 *   occ_time        f64[N]     out or NULL (zeroed here, :54) -- only if record
 *   trace           int32[2*hops] out or NULL: (from,to) per hop
 * returns simulated time (starts from 0, :55).
 */
double kmc_oracle_py(int N, int P, double nu, double kT, double I_0, double R,
                     uint8_t *occupation, const double *distances,
                     const double *E_constant, double *site_energies,
                     const double *transitions_constant, int64_t *electrode_occ,
                     int64_t hops, int record, const double *u, uint64_t seed,
                     double *traffic, double *occ_time, int32_t *trace)
{
    const int S = N + P;
    double *transitions = (double *)malloc(sizeof(double) * S * S);
    double *problist = (double *)malloc(sizeof(double) * S * S);
    xo_t g; xo_seed(&g, seed);
    double time = 0.0;
    if (record && traffic) memset(traffic, 0, sizeof(double) * S * S);
    if (record && occ_time) memset(occ_time, 0, sizeof(double) * N);

    for (int64_t hop = 0; hop < hops; hop++) {
        /* :58-66 site energies from scratch */
        for (int i = 0; i < N; i++) {
            double acc = 0.0;
            for (int j = 0; j < N; j++)
                if (j != i) acc += (double)(1 - (int)occupation[j]) / distances[i * S + j];
            site_energies[i] = E_constant[i] + (-I_0 * R * acc);
        }
        /* :68-87 Miller-Abrahams rates */
        for (int i = 0; i < S; i++)
            for (int j = 0; j < S; j++) {
                double t;
                if (!transition_possible(i, j, N, occupation)) t = 0.0;
                else {
                    double dE;
                    if (i < N && j < N)
                        dE = site_energies[j] - site_energies[i] - I_0 * R / distances[i * S + j];
                    else
                        dE = site_energies[j] - site_energies[i];
                    t = (dE > 0) ? nu * exp(-dE / kT) : nu;
                }
                transitions[i * S + j] = transitions_constant[i * S + j] * t;
            }
        /* :91-97 cumulative list */
        problist[0] = transitions[0];
        for (int k = 1; k < S * S; k++) problist[k] = transitions[k] + problist[k - 1];
        /* :100 dwell time; one uniform consumed */
        const double total = problist[S * S - 1];
        const double u1 = u ? u[2 * hop] : xo_double(&g);
        const double hop_time = -log(1.0 - u1) * (1.0 / total);
        /* :103-110 normalised pick; second uniform */
        const double u2 = u ? u[2 * hop + 1] : xo_double(&g);
        int event = 0; /* no hit (total==0 -> nan list): int(u2/S)=0, int(u2%S)=0 */
        for (int k = 0; k < S * S; k++)
            if (problist[k] / total >= u2) { event = k; break; }
        const int from = event / S, to = event % S;
        /* :115-123 perform hop */
        if (from < N) occupation[from] = 0; else electrode_occ[from - N] -= 1;
        if (to < N) occupation[to] = 1; else electrode_occ[to - N] += 1;
        /* :126-130 record (post-hop occupation) */
        if (record) {
            if (traffic) traffic[from * S + to] += 1.0;
            if (occ_time)
                for (int i = 0; i < N; i++) if (occupation[i]) occ_time[i] += hop_time;
        }
        if (trace) { trace[2 * hop] = from; trace[2 * hop + 1] = to; }
        time += hop_time; /* :133 */
    }
    free(transitions); free(problist);
    return time;
}

/* ===================================================================== (B) */
typedef struct { int from, to; float rate; } transition_t;

/* simulation.go:29-38 */
static inline uint64_t get_key(const uint8_t *occ, int N) {
    uint64_t r = 0;
    for (int i = 0; i < N; i++) { r = r << 1; if (occ[i]) r += 1; }
    return r;
}

/* open-addressing map uint64 -> (count:uint16, list index) standing in for Go's
 * allProbs / countProbs maps (simulation.go:222-223, 351-352). */
typedef struct { uint64_t key; int32_t list; uint16_t count; uint8_t used; } slot_t;
typedef struct { slot_t *slots; size_t cap, n; } map_t;
static void map_init(map_t *m, size_t cap) { m->cap = cap; m->n = 0; m->slots = (slot_t *)calloc(cap, sizeof(slot_t)); }
static inline uint64_t hash64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
static slot_t *map_find(map_t *m, uint64_t key, int insert);
static void map_grow(map_t *m) {
    map_t n; map_init(&n, m->cap * 2);
    for (size_t i = 0; i < m->cap; i++) if (m->slots[i].used) { slot_t *s = map_find(&n, m->slots[i].key, 1); s->list = m->slots[i].list; s->count = m->slots[i].count; s->used = m->slots[i].used; }
    free(m->slots); *m = n;
}
static slot_t *map_find(map_t *m, uint64_t key, int insert) {
    if (insert && (m->n + 1) * 2 > m->cap) map_grow(m);
    size_t i = hash64(key) & (m->cap - 1);
    for (;;) {
        slot_t *s = &m->slots[i];
        if (!s->used) {
            if (!insert) return NULL;
            s->used = 1; s->key = key; s->list = -1; s->count = 0; m->n++;
            return s;
        }
        if (s->key == key) return s;
        i = (i + 1) & (m->cap - 1);
    }
}

/* simulation.go:163-188: first index with probList[i] >= r (lower bound; the
 * halving-step walk of the reference terminates exactly there because the list
 * is non-decreasing).  r==0 selects index 0 even if its rate is 0. */
static inline int get_random_event(const float *probList, int len, float eventRand) {
    int i = len / 2, e_step = len / 2;
    for (;;) {
        if (e_step >= 2) e_step = e_step / 2;
        if (probList[i] < eventRand) { i += e_step; if (i >= len) i = len - 1; }
        else if (i > 0 && probList[i - 1] >= eventRand) { i -= e_step; if (i < 0) i = 0; }
        else return i;
    }
}

/* simulation.go:226-234 / :378-386 : fp32 energies from scratch */
static void go_energies_scratch(int N, int S, const float *d32, const uint8_t *occ,
                                const float *E32, float I_0, float R, float *se)
{
    for (int i = 0; i < N; i++) {
        float acc = 0.0f;
        for (int j = 0; j < N; j++)
            if (j != i && !occ[j]) acc += 1.0f / d32[i * S + j];
        se[i] = E32[i] - I_0 * R * acc;
    }
}

/* simulation.go:58-80 */
static void go_calc_transition_list(transition_t *tr, int L, const float *d32, const uint8_t *occ,
                                    const float *se, float R, float I_0, float kT, float nu,
                                    int N, int S, const float *tc32)
{
    for (int k = 0; k < L; k++) {
        const int from = tr[k].from, to = tr[k].to;
        if (!transition_possible(from, to, N, occ)) { tr[k].rate = 0.0f; continue; }
        float dE;
        if (from < N && to < N) dE = se[to] - se[from] - I_0 * R / d32[from * S + to];
        else dE = se[to] - se[from];
        float rate;
        if (dE > 0) rate = nu * (float)exp((double)(-dE / kT));
        else rate = nu;
        rate *= tc32[from * S + to];
        tr[k].rate = rate;
    }
}

/* simulation.go:107-130 */
static void go_make_jump(uint8_t *occ, double *eo, float *se, const float *d32,
                         float R, float I_0, int N, int S, int from, int to)
{
    if (from < N) {
        occ[from] = 0;
        for (int j = 0; j < N; j++) if (j != from) se[j] -= I_0 * R * (1.0f / d32[j * S + from]);
    } else eo[from - N] -= 1.0;
    if (to < N) {
        occ[to] = 1;
        for (int j = 0; j < N; j++) if (j != to) se[j] += I_0 * R * (1.0f / d32[j * S + to]);
    } else eo[to - N] += 1.0;
}

/*
 * kmc_oracle_go -- restatement of simulate (variant 0, simulation.go:194-325)
 * and simulateRecordPlus (variant 1, :327-432) behind the narrowing the
 * wrappers apply (simulationWrapper.go:83-169, 250-272).
 *
 *   occupation_in  f64[N] or NULL.  NULL = all-empty start, which is what every
 *                  single-run export does (simulationWrapper.go:90,105,134-141,
 *                  156-163); non-NULL = honoured as ">0" (channelSimulateRecord,
 *                  :253-260).
 *   electrode_occ  f64[P] out (zeroed first, simulation.go:236-238 / :355-357)
 *   site_energies  f64[S] in ([N:] electrode energies).  NOT written back: the
 *                  wrappers hand the loop a float32 copy (toFloat32).
 *   use_cache      record_problist flag (state cache, simulation.go:254-295 /
 *                  :367-412).  0 for wrapperSimulate / wrapperSimulatePruned.
 *   cut            transition_cut_constant (prune threshold), 0 keeps tc>0 pairs.
 *   e,u            injected stream (f64 Exp(1) variates, f32 uniforms) or NULL.
 *   occupation_out uint8[N] or NULL: final occupation.
 *   trace          int32[2*hops] or NULL.
 *   se_out         f32[S] or NULL: final fp32 site energies.
 *   n_miss         int64* or NULL: number of rate-list evaluations (cache misses).
 */
double kmc_oracle_go(int variant, int N, int P, double nu64, double kT64, double I_064, double R64,
                     const double *occupation_in, const double *distances, const double *E_constant,
                     const double *transitions_constant, double *electrode_occ,
                     const double *site_energies, int64_t hops, int use_cache, int record,
                     double cut64, const double *e, const float *u, uint64_t seed,
                     double *traffic, double *average_occupation,
                     uint8_t *occupation_out, int32_t *trace, float *se_out, int64_t *n_miss)
{
    const int S = N + P;
    const float nu = (float)nu64, kT = (float)kT64, I_0 = (float)I_064, R = (float)R64;
    const float cut = (float)cut64;
    float *d32 = (float *)malloc(sizeof(float) * S * S);
    float *tc32 = (float *)malloc(sizeof(float) * S * S);
    float *E32 = (float *)malloc(sizeof(float) * (N > 0 ? N : 1));
    float *se = (float *)malloc(sizeof(float) * S);
    uint8_t *occ = (uint8_t *)calloc(N > 0 ? N : 1, 1);
    /* simulationWrapper.go:37-56 narrowing */
    for (int k = 0; k < S * S; k++) { d32[k] = (float)distances[k]; tc32[k] = (float)transitions_constant[k]; }
    for (int i = 0; i < N; i++) E32[i] = (float)E_constant[i];
    for (int i = 0; i < S; i++) se[i] = (float)site_energies[i];
    if (occupation_in) for (int i = 0; i < N; i++) occ[i] = occupation_in[i] > 0;

    /* simulation.go:199-215 transition list */
    transition_t *tr = (transition_t *)malloc(sizeof(transition_t) * S * S);
    int L = 0;
    float largest = 0.0f;
    for (int k = 0; k < S * S; k++) if (largest < tc32[k]) largest = tc32[k];
    for (int i = 0; i < S; i++)
        for (int j = 0; j < S; j++)
            if (tc32[i * S + j] > cut * largest) { tr[L].from = i; tr[L].to = j; tr[L].rate = 0; L++; }

    if (variant == 0) go_energies_scratch(N, S, d32, occ, E32, I_0, R, se); /* :226-234 */
    for (int i = 0; i < P; i++) electrode_occ[i] = 0.0;
    double time = 0.0;

    /* state cache */
    map_t map; map.slots = NULL;
    float *store = NULL; size_t store_cap = 0, store_n = 0;
    uint64_t countStorage = 0, reuseThresholdIncrease = 100000, allowedSaves = L ? (uint64_t)(150000000 / L) : 0;
    uint16_t reuseThreshold = 1;
    if (use_cache) map_init(&map, 1024);
    float *scratch = (float *)malloc(sizeof(float) * (L > 0 ? L : 1));
    xo_t g; xo_seed(&g, seed);
    int64_t misses = 0;

    for (int64_t hop = 0; hop < hops; hop++) {
        const float *probList = NULL;
        slot_t *slot = NULL;
        if (use_cache) {
            slot = map_find(&map, get_key(occ, N), 1);
            if (slot->list >= 0) probList = store + (size_t)slot->list * L;
        }
        if (!probList) {
            misses++;
            if (variant == 1) go_energies_scratch(N, S, d32, occ, E32, I_0, R, se); /* :378-386 */
            go_calc_transition_list(tr, L, d32, occ, se, R, I_0, kT, nu, N, S, tc32);
            float *pl = scratch;
            int do_store = 0;
            if (use_cache) {
                /* simulation.go:278-295 / :400-412 */
                if (slot->count > 0 || slot->used == 2) {
                    const uint16_t val = slot->count;
                    slot->count++;
                    if (variant == 0) {
                        if (val >= reuseThreshold) {
                            do_store = 1; countStorage++;
                            if (countStorage > reuseThresholdIncrease) { reuseThreshold++; reuseThresholdIncrease += 100000; }
                        }
                    } else if (countStorage < allowedSaves) { do_store = 1; countStorage++; }
                } else { slot->count = 1; slot->used = 2; }
            }
            if (do_store) {
                if (store_n == store_cap) {
                    store_cap = store_cap ? store_cap * 2 : 64;
                    store = (float *)realloc(store, sizeof(float) * store_cap * (size_t)L);
                }
                slot->list = (int32_t)store_n;
                pl = store + store_n * (size_t)L;
                store_n++;
            }
            for (int k = 0; k < L; k++) pl[k] = (k == 0) ? tr[0].rate : pl[k - 1] + tr[k].rate; /* :270-276 */
            probList = pl;
        }
        const float total = probList[L - 1];
        const double ek = e ? e[hop] : xo_exp(&g);
        const double time_step = ek / (double)total; /* :297 */
        time += time_step;
        const float uk = u ? u[hop] : xo_float(&g);
        const int event = get_random_event(probList, L, uk * total); /* :299, :164 */
        const int from = tr[event].from, to = tr[event].to;
        if (trace) { trace[2 * hop] = from; trace[2 * hop + 1] = to; }
        if (variant == 0) {
            if (record) { /* :309-317 (pre-hop occupation, antisymmetric traffic) */
                if (traffic) { traffic[from * S + to] += 1.0; traffic[to * S + from] -= 1.0; }
                if (average_occupation)
                    for (int i = 0; i < N; i++) if (occ[i]) average_occupation[i] += time_step;
            }
            go_make_jump(occ, electrode_occ, se, d32, R, I_0, N, S, from, to);
        } else { /* :420-429 */
            if (from < N) occ[from] = 0; else electrode_occ[from - N] -= 1.0;
            if (to < N) occ[to] = 1; else electrode_occ[to - N] += 1.0;
        }
    }
    if (occupation_out) memcpy(occupation_out, occ, N);
    if (se_out) memcpy(se_out, se, sizeof(float) * S);
    if (n_miss) *n_miss = misses;
    if (use_cache) free(map.slots);
    free(store); free(scratch); free(tr); free(d32); free(tc32); free(E32); free(se); free(occ);
    return time;
}

/* One evaluation of the fp32 energies + rate list for a GIVEN occupation
 * (simulation.go:226-234 then :58-80, cut=0 list).  Used by the 1e-6-relative
 * rate/energy parity tests.  rates: f32[S*S] dense row-major (0 where the pair
 * is not in the list or not allowed); se: f32[S]. */
void kmc_oracle_go_rates(int N, int P, double nu64, double kT64, double I_064, double R64,
                         const uint8_t *occupation, const double *distances, const double *E_constant,
                         const double *transitions_constant, const double *site_energies,
                         float *se_out, float *rates_out)
{
    const int S = N + P;
    const float nu = (float)nu64, kT = (float)kT64, I_0 = (float)I_064, R = (float)R64;
    float *d32 = (float *)malloc(sizeof(float) * S * S);
    float *tc32 = (float *)malloc(sizeof(float) * S * S);
    float *E32 = (float *)malloc(sizeof(float) * (N > 0 ? N : 1));
    transition_t *tr = (transition_t *)malloc(sizeof(transition_t) * S * S);
    for (int k = 0; k < S * S; k++) { d32[k] = (float)distances[k]; tc32[k] = (float)transitions_constant[k]; }
    for (int i = 0; i < N; i++) E32[i] = (float)E_constant[i];
    for (int i = 0; i < S; i++) se_out[i] = (float)site_energies[i];
    go_energies_scratch(N, S, d32, occupation, E32, I_0, R, se_out);
    int L = 0;
    for (int i = 0; i < S; i++) for (int j = 0; j < S; j++) if (tc32[i * S + j] > 0.0f) { tr[L].from = i; tr[L].to = j; L++; }
    go_calc_transition_list(tr, L, d32, occupation, se_out, R, I_0, kT, nu, N, S, tc32);
    memset(rates_out, 0, sizeof(float) * S * S);
    for (int k = 0; k < L; k++) rates_out[tr[k].from * S + tr[k].to] = tr[k].rate;
    free(d32); free(tc32); free(E32); free(tr);
}

/* fp64 energies + rate matrix for a given occupation, numba semantics
 * (kmc_dopant_networks.py:58-87). */
void kmc_oracle_py_rates(int N, int P, double nu, double kT, double I_0, double R,
                         const uint8_t *occupation, const double *distances, const double *E_constant,
                         const double *transitions_constant, const double *site_energies,
                         double *se_out, double *rates_out)
{
    const int S = N + P;
    for (int i = 0; i < S; i++) se_out[i] = site_energies[i];
    for (int i = 0; i < N; i++) {
        double acc = 0.0;
        for (int j = 0; j < N; j++) if (j != i) acc += (double)(1 - (int)occupation[j]) / distances[i * S + j];
        se_out[i] = E_constant[i] + (-I_0 * R * acc);
    }
    for (int i = 0; i < S; i++)
        for (int j = 0; j < S; j++) {
            double t = 0.0;
            if (transition_possible(i, j, N, occupation)) {
                double dE = (i < N && j < N) ? se_out[j] - se_out[i] - I_0 * R / distances[i * S + j]
                                             : se_out[j] - se_out[i];
                t = (dE > 0) ? nu * exp(-dE / kT) : nu;
            }
            rates_out[i * S + j] = transitions_constant[i * S + j] * t;
        }
}

/*
 * Ensemble runners -- B independent members of ONE layout on host threads
 * (pthreads, members handed out one at a time from an atomic counter): the
 * shape of parallelSimulations (simulationWrapper.go:274-316: one goroutine per
 * simulation running simulateRecordPlus with record_problist=true, input
 * occupation honoured).  CPU-baseline timing and statistical checks.
 */
typedef struct {
    int semantics;          /* 0 = go, 1 = py */
    int variant, use_cache, N, P;
    int64_t B, hops;
    double nu, I_0, R;
    const double *kT, *occupation0, *distances, *E_constant, *transitions_constant, *electrode_v;
    uint64_t seed0;
    double *time_out, *eo_out_f64;
    int64_t *eo_out_i64;
    atomic_llong next;
} ens_job_t;

static void *ens_worker(void *arg)
{
    ens_job_t *J = (ens_job_t *)arg;
    const int N = J->N, P = J->P, S = N + P;
    double *se = (double *)calloc(S, sizeof(double));
    uint8_t *occ = (uint8_t *)calloc(N > 0 ? N : 1, 1);
    for (;;) {
        const int64_t m = atomic_fetch_add(&J->next, 1);
        if (m >= J->B) break;
        for (int p = 0; p < P; p++) se[N + p] = J->electrode_v[m * P + p];
        if (J->semantics == 0) {
            J->time_out[m] = kmc_oracle_go(J->variant, N, P, J->nu, J->kT[m], J->I_0, J->R, J->occupation0,
                                           J->distances, J->E_constant + m * N, J->transitions_constant,
                                           J->eo_out_f64 + m * P, se, J->hops, J->use_cache, 0, 0.0,
                                           NULL, NULL, J->seed0 + (uint64_t)m, NULL, NULL, NULL, NULL, NULL, NULL);
        } else {
            for (int i = 0; i < N; i++) occ[i] = J->occupation0 ? (J->occupation0[i] > 0) : 0;
            for (int p = 0; p < P; p++) J->eo_out_i64[m * P + p] = 0;
            J->time_out[m] = kmc_oracle_py(N, P, J->nu, J->kT[m], J->I_0, J->R, occ, J->distances,
                                           J->E_constant + m * N, se, J->transitions_constant,
                                           J->eo_out_i64 + m * P, J->hops, 0, NULL,
                                           J->seed0 + (uint64_t)m, NULL, NULL, NULL);
        }
    }
    free(se); free(occ);
    return NULL;
}

static int ens_run(ens_job_t *J, int nthreads)
{
    if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads < 1) nthreads = 1;
    if ((int64_t)nthreads > J->B) nthreads = (int)(J->B > 0 ? J->B : 1);
    atomic_init(&J->next, 0);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, ens_worker, J);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    return nthreads;
}

/*   E_constant f64[B,N]; electrode_v f64[B,P]; kT f64[B]; occupation0 f64[N] or NULL
 *   out: time f64[B], electrode_occ f64[B,P].  returns the number of threads used. */
int kmc_oracle_go_ensemble(int variant, int use_cache, int N, int P, int64_t B,
                           double nu, const double *kT, double I_0, double R,
                           const double *occupation0, const double *distances,
                           const double *E_constant, const double *transitions_constant,
                           const double *electrode_v, int64_t hops, uint64_t seed0,
                           int nthreads, double *time_out, double *electrode_occ_out)
{
    ens_job_t J;
    memset(&J, 0, sizeof(J));
    J.semantics = 0; J.variant = variant; J.use_cache = use_cache; J.N = N; J.P = P; J.B = B; J.hops = hops;
    J.nu = nu; J.I_0 = I_0; J.R = R; J.kT = kT; J.occupation0 = occupation0; J.distances = distances;
    J.E_constant = E_constant; J.transitions_constant = transitions_constant; J.electrode_v = electrode_v;
    J.seed0 = seed0; J.time_out = time_out; J.eo_out_f64 = electrode_occ_out;
    return ens_run(&J, nthreads);
}

/* Same for the numba semantics (CPU baseline of python_simulation); electrode
 * tallies are int64 there. */
int kmc_oracle_py_ensemble(int N, int P, int64_t B, double nu, const double *kT, double I_0, double R,
                           const double *occupation0, const double *distances,
                           const double *E_constant, const double *transitions_constant,
                           const double *electrode_v, int64_t hops, uint64_t seed0,
                           int nthreads, double *time_out, int64_t *electrode_occ_out)
{
    ens_job_t J;
    memset(&J, 0, sizeof(J));
    J.semantics = 1; J.N = N; J.P = P; J.B = B; J.hops = hops;
    J.nu = nu; J.I_0 = I_0; J.R = R; J.kT = kT; J.occupation0 = occupation0; J.distances = distances;
    J.E_constant = E_constant; J.transitions_constant = transitions_constant; J.electrode_v = electrode_v;
    J.seed0 = seed0; J.time_out = time_out; J.eo_out_i64 = electrode_occ_out;
    return ens_run(&J, nthreads);
}

/* ===================================================================== (C) */
/*
 * kmc_oracle_prob -- restatement of the mean-field "probability" solver probSimulate
 * (goSimulation/probabilitySimulation.go:53-157, calcProbTransitions :8-38, probTransitionPossible :40-50)
 * behind wrapperSimulateProbability (simulationWrapper.go:218-233: fp64 tables, occupation forced to 0.5).
 * Deterministic (no random numbers), fp64 throughout, same loop order as the reference.
 *   occupation   f64[N] out (fractional occupations after `hops` relaxation steps; starts at 0.5)
 *   electrode_occ f64[P] out;  site_energies f64[S] in ([N:]) / out ([:N])
 *   traffic f64[S*S], average_occupation f64[N]: accumulated when record (caller zeroes)
 * returns the accumulated time.
 */
double kmc_oracle_prob(int NSites, int NElectrodes, double nu, double kT, double I_0, double R,
                       double *occupation, const double *distances, const double *E_constant,
                       const double *transitions_constant, double *electrode_occupation,
                       double *site_energies, int64_t hops, int record, double *traffic,
                       double *average_occupation)
{
    const int N = NSites + NElectrodes;
    double *transitions = (double *)calloc((size_t)N * N, sizeof(double));
    double *difference = (double *)calloc(NSites > 0 ? NSites : 1, sizeof(double));
    for (int j = 0; j < NSites; j++) occupation[j] = 0.5;             /* simulationWrapper.go:226-228 */
    for (int i = 0; i < NElectrodes; i++) electrode_occupation[i] = 0.0;
    double time = 0.0, tot_rates = 0.0;
    for (int64_t hop = 0; hop < hops; hop++) {
        for (int i = 0; i < NSites; i++) {                            /* :84-93 */
            difference[i] = 0;
            double acc = 0.0;
            for (int j = 0; j < NSites; j++)
                if (j != i) acc += (1 - occupation[j]) / distances[i * N + j];
            site_energies[i] = E_constant[i] - I_0 * R * acc;
        }
        tot_rates = 0.0;                                              /* calcProbTransitions :8-38 */
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) {
                double base;
                if (i >= NSites && j >= NSites) base = 0;
                else if (i >= NSites) base = 1 - occupation[j];
                else if (j >= NSites) base = occupation[i];
                else base = (1 - occupation[j]) * occupation[i];
                double dE;
                if (i < NSites && j < NSites) dE = site_energies[j] - site_energies[i] - I_0 * R / distances[i * N + j];
                else dE = site_energies[j] - site_energies[i];
                double t = (dE > 0) ? base * nu * exp(-dE / kT) : base * nu;
                t *= transitions_constant[i * N + j];
                transitions[i * N + j] = t;
                tot_rates += t;
                if (i < NSites) difference[i] -= t;
                if (j < NSites) difference[j] += t;
            }
        double max_change = 1.0;                                      /* :99-114 */
        for (int i = 0; i < NSites; i++) {
            const double newVal = occupation[i] + difference[i] / tot_rates;
            if (newVal < 0) {
                const double req = occupation[i] / -(difference[i] / tot_rates);
                if (req < max_change) max_change = req;
            }
            if (newVal > 1) {
                const double req = (1 - occupation[i]) / (difference[i] / tot_rates);
                if (req < max_change) max_change = req;
            }
        }
        const double time_step = 0.98 * max_change / tot_rates;
        time += time_step;
        if (record && average_occupation)
            for (int i = 0; i < NSites; i++) average_occupation[i] += occupation[i] * time_step;
        for (int i = 0; i < N; i++)                                   /* :123-145 */
            for (int j = 0; j < N; j++) {
                if (i >= NSites && j >= NSites) break;
                const double rate = transitions[i * N + j] * max_change / tot_rates;
                if (i < NSites) occupation[i] -= rate; else electrode_occupation[i - NSites] -= rate;
                if (j < NSites) occupation[j] += rate; else electrode_occupation[j - NSites] += rate;
                if (record && traffic) { traffic[i * N + j] += rate; traffic[j * N + i] -= rate; }
            }
        for (int i = 0; i < NSites; i++) {                            /* :146-155 */
            if (occupation[i] < 0) occupation[i] = 0;
            if (occupation[i] > 1) occupation[i] = 1;
        }
    }
    free(transitions); free(difference);
    return time;
}
