/*
 * kmc_b200.h -- C ABI of libkmcb200.so, the B200 (sm_100a) replacement of
 * kmc_dn's goSimulation/libSimulation.so for the KMC hop loop.
 *
 * Part 1 mirrors, symbol for symbol, the cgo exports the reference's ctypes
 * bindings load (goSimulation/pythonBind.py:49-90,
 * goSimulation/parrallelSimulationBind.py:50-66).  Part 2 is the additive lean
 * ensemble API (raw pointers, no GoSlice boxing) used by kmc_dn_b200's host
 * class and bench.py.
 *
 * No torch types, no C++ types: plain pointers and sizes only.
 * All file:line citations are relative to the reference tree (MUTUEL/kmc_dn).
 */
#ifndef KMC_B200_H
#define KMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * Part 1 -- drop-in exports of libSimulation.so
 * ---------------------------------------------------------------------- */

/* Go slice header as the reference's ctypes code builds it
 * (goSimulation/pythonBind.py:28-30), PASSED BY VALUE. */
typedef struct {
    double *data;
    long long len;
    long long cap;
} GoSlice;

/*
 * Single-trajectory exports.  Signature of every one of them
 * (goSimulation/simulationWrapper.go:83,127,149):
 *   NSites = number of acceptors, NElectrodes = number of electrodes,
 *   occupation[N] (ignored: every single-run export starts from the all-empty
 *   state, simulationWrapper.go:90,134-141,156-163), distances[S*S],
 *   E_constant[N], transitions_constant[S*S], electrode_occupation[P] (out,
 *   zeroed first, simulation.go:236-238), site_energies[S] ([N:] = electrode
 *   energies, in), hops, record, traffic[S*S] (out), average_occupation[N] (out).
 *   Returns the simulated time.
 *
 * ABI quirk kept on purpose: the reference's Python caller passes FIVE doubles
 * (nu,kT,I_0,R,time) and a 32-bit hops (pythonBind.py:65-72) while cgo declares
 * four doubles and a 64-bit int.  Declaring the trailing `time_unused` double
 * and `int hops` accepts both callers on x86-64 SysV (XMM4 is simply unused by
 * a 4-double caller).
 */
double wrapperSimulate(long long NSites, long long NElectrodes, double nu, double kT, double I_0,
                       double R, double time_unused, GoSlice occupation, GoSlice distances,
                       GoSlice E_constant, GoSlice transitions_constant, GoSlice electrode_occupation,
                       GoSlice site_energies, int hops, unsigned char record, GoSlice traffic,
                       GoSlice average_occupation);                /* simulationWrapper.go:83-96  */
double wrapperSimulateRecord(long long NSites, long long NElectrodes, double nu, double kT, double I_0,
                             double R, double time_unused, GoSlice occupation, GoSlice distances,
                             GoSlice E_constant, GoSlice transitions_constant,
                             GoSlice electrode_occupation, GoSlice site_energies, int hops,
                             unsigned char record, GoSlice traffic,
                             GoSlice average_occupation);          /* simulationWrapper.go:127-147 */
double wrapperSimulateRecordPlus(long long NSites, long long NElectrodes, double nu, double kT,
                                 double I_0, double R, double time_unused, GoSlice occupation,
                                 GoSlice distances, GoSlice E_constant, GoSlice transitions_constant,
                                 GoSlice electrode_occupation, GoSlice site_energies, int hops,
                                 unsigned char record, GoSlice traffic,
                                 GoSlice average_occupation);      /* simulationWrapper.go:149-169 */
double wrapperSimulateProbability(long long NSites, long long NElectrodes, double nu, double kT,
                                  double I_0, double R, double time_unused, GoSlice occupation,
                                  GoSlice distances, GoSlice E_constant, GoSlice transitions_constant,
                                  GoSlice electrode_occupation, GoSlice site_energies, int hops,
                                  unsigned char record, GoSlice traffic,
                                  GoSlice average_occupation);     /* simulationWrapper.go:218-233: mean-field
                                     solver; WRITES the fractional occupations into `occupation` and the acceptor
                                     energies into `site_energies`, as the Go code does                      */
/* extra leading prune_threshold (simulationWrapper.go:98-110, pythonBind.py:73-79) */
double wrapperSimulatePruned(long long NSites, long long NElectrodes, double prune_threshold, double nu,
                             double kT, double I_0, double R, double time_unused, GoSlice occupation,
                             GoSlice distances, GoSlice E_constant, GoSlice transitions_constant,
                             GoSlice electrode_occupation, GoSlice site_energies, int hops,
                             unsigned char record, GoSlice traffic, GoSlice average_occupation);

/*
 * Batched export (simulationWrapper.go:274-316; caller
 * parrallelSimulationBind.py:50-66).  14 slices; per-simulation scalars are
 * double arrays of length B; occupation / E_constant concatenated sum(N_i);
 * electrode_occupation sum(P_i); site_energies sum(S_i); distances /
 * transitions_constant concatenated sum(S_i^2) row-major blocks.  Input
 * occupation IS honoured here (simulationWrapper.go:253-260).  Writes
 * electrode_occupation and time in place, returns 0.
 * Simulations with byte-identical (N,P,nu,I_0,R,distances,transitions_constant)
 * share one device layout and run as one ensemble launch.
 */
long long parallelSimulations(GoSlice NSites, GoSlice NElectrodes, GoSlice nu, GoSlice kT, GoSlice I_0,
                              GoSlice R, GoSlice occupation, GoSlice distances, GoSlice E_constant,
                              GoSlice transitions_constant, GoSlice electrode_occupation, GoSlice hops,
                              GoSlice time, GoSlice site_energies);

/* ------------------------------------------------------------------------
 * Part 2 -- lean ensemble API (additive; not in the reference)
 * ---------------------------------------------------------------------- */

typedef struct kmcb200_layout kmcb200_layout; /* device-resident tables of one dopant layout */

/* hop-loop arithmetic */
enum {
    KMCB200_MODE_FAST = 0,          /* production: fp32 rates (ex2.approx), fp64 incremental energies,
                                       fp64 cumulative rates + time, Philox4x32-10                      */
    KMCB200_MODE_GO_SIMULATE = 1,   /* replay: op-for-op simulate,           simulation.go:194-325     */
    KMCB200_MODE_GO_RECORDPLUS = 2, /* replay: op-for-op simulateRecordPlus, simulation.go:327-432     */
    KMCB200_MODE_PY = 3,            /* replay: op-for-op numba loop, kmc_dopant_networks.py:33-135     */
    KMCB200_MODE_FAST_REFORDER = 4, /* production arithmetic with the reference's row-major event order:
                                       follows the Go loop hop for hop under an injected stream           */
    KMCB200_MODE_PROB = 5           /* mean-field pre-screen probSimulate, probabilitySimulation.go:53-157
                                       (deterministic, fp64; `hops` = relaxation steps; occupation starts
                                       at 0.5; results in prob_occupation / prob_electrode_occ / time)    */
};

enum {
    KMCB200_FLAG_DEVICE_PTRS = 1, /* every data pointer in the args is a device pointer on layout's GPU;
                                     nothing is copied and the call returns after enqueueing on `stream` */
    KMCB200_FLAG_NO_MEMO = 2,     /* MODE_FAST: disable the per-warp state memoisation (results are
                                     bit-identical either way; for testing and profiling)                */
    KMCB200_FLAG_LANES = 4,       /* MODE_FAST, N <= 31, fewer than 2^31 hops: force the thread-per-trajectory kernel
                                     (default: chosen for ensembles of >= 12288 members)                     */
    KMCB200_FLAG_NO_LANES = 8,    /* MODE_FAST: never use the thread-per-trajectory kernel                 */
    KMCB200_FLAG_SOLO = 16,       /* MODE_FAST, N <= 31, fewer than 2^31 hops, no record / trace / injected-stream outputs:
                                     force the latency kernel (one warp per trajectory, the visited states as a graph in
                                     shared memory; default: chosen for ensembles of at most 8 members per SM)         */
    KMCB200_FLAG_NO_SOLO = 32     /* MODE_FAST: never use the latency kernel                                       */
};

typedef struct {
    int64_t B;        /* ensemble members                                                              */
    int64_t hops;     /* recorded hops per member                                                      */
    int64_t prehops;  /* equilibration hops before tallies are reset (kmc_dopant_networks.py:580-585)  */
    int32_t mode;     /* KMCB200_MODE_*                                                                */
    int32_t flags;    /* KMCB200_FLAG_*                                                                */
    /* --- per-member inputs.  Either E_constant, or basis + electrode_v (superposition mat-vec on
     *     device: E_constant[m,i] = basis[P,i] + sum_p electrode_v[m,p]*basis[p,i]). */
    const double *E_constant;   /* [B,N] or NULL                                                       */
    const double *basis;        /* [P+1,N] or NULL                                                     */
    const double *electrode_v;  /* [B,P]  electrode energies = site_energies[N:]                       */
    const double *kT;           /* [B]                                                                 */
    const uint8_t *occupation0; /* [B,N] initial occupation, or NULL = all empty                       */
    uint64_t seed;              /* member m draws from Philox stream (seed, member_index0 + m)         */
    uint64_t member_index0;     /* global index of member 0 (multi-GPU shards keep global numbering)   */
    /* --- injected random stream for replay (all NULL = on-device Philox).
     *     go/fast modes: e [B,prehops+hops] f64 Exp(1) variates + u [B,prehops+hops] f32 uniforms.
     *     py mode:       u64 [B,2*(prehops+hops)] f64 uniforms (dwell, pick). */
    const double *stream_e;
    const float *stream_u;
    const double *stream_u64;
    /* --- outputs */
    double *time;               /* [B]                                                                 */
    int64_t *electrode_occ;     /* [B,P] net holes into each electrode                                 */
    uint8_t *occupation_out;    /* [B,N] or NULL                                                       */
    double *site_energies_out;  /* [B,S] or NULL (final energies as the loop holds them)               */
    double *avg_occupation;     /* [B,N] or NULL: un-normalised occupied time (record)                 */
    double *traffic;            /* [B,S,S] or NULL (record)                                            */
    int32_t *trace;             /* [B,hops,2] or NULL: (from,to) of every recorded hop                 */
    int64_t *misses;            /* [B] or NULL: hops whose rate structure had to be evaluated (MODE_FAST:
                                   state-cache misses; equals prehops+hops with KMCB200_FLAG_NO_MEMO)     */
    double *prob_occupation;    /* [B,N] or NULL: fractional occupations (MODE_PROB)                   */
    double *prob_electrode_occ; /* [B,P] or NULL: fractional electrode tallies (MODE_PROB)             */
    void *stream;               /* cudaStream_t, NULL = default stream                                 */
} kmcb200_ensemble_args;

int kmcb200_device_count(void);
const char *kmcb200_last_error(void);
const char *kmcb200_version(void);
/* sizeof(kmcb200_ensemble_args) as compiled, so that foreign-language bindings can check their mirror. */
int kmcb200_sizeof_ensemble_args(void);

/* Seed used by the Part-1 exports (the reference never seeds Go's global generator,
 * simulation.go:164,297).  Each export call consumes one stream index. */
void kmcb200_set_seed(uint64_t seed);

/* Tables are narrowed to float32 exactly as the cgo wrappers do
 * (simulationWrapper.go:37-56); prune_threshold as in simulation.go:200-215. */
kmcb200_layout *kmcb200_layout_create(int device, int N, int P, const double *distances,
                                      const double *transitions_constant, double nu, double I_0,
                                      double R, double prune_threshold);
void kmcb200_layout_destroy(kmcb200_layout *layout);

/* Runs the ensemble on the layout's device.  Returns 0 on success, non-zero on error
 * (message via kmcb200_last_error).  Host-pointer calls are synchronous.  Calls with KMCB200_FLAG_DEVICE_PTRS
 * return after enqueueing; ONE launch is in flight per layout: the next call on the same layout -- from any
 * stream -- waits on the device for the previous one (work queue, workspace and state tables belong to the
 * layout).  For concurrent launches use one layout per stream. */
int kmcb200_run_ensemble(kmcb200_layout *layout, const kmcb200_ensemble_args *args);

/* The same call spread over several GPUs of one box from ONE process (the ensemble shards
 * trivially; the reference's counterpart is the goroutine fan-out of parallelSimulations,
 * simulationWrapper.go:280-314): `layouts` are n_layouts copies of one layout created on different devices; the
 * members are cut into contiguous blocks, one host thread per device.  Host pointers only, stream must be NULL.
 * Results are identical to a single-device call (streams are numbered by global member index). */
int kmcb200_run_ensemble_multi(kmcb200_layout *const *layouts, int n_layouts, const kmcb200_ensemble_args *args);

/* Per-voltage-vector statistics of the currents on the device (the reduction the reference's consumers do
 * on the host: voltage_search.py:160-185, validate_tests.py:80-135).  All DEVICE pointers on `device`, asynchronous on
 * `stream`.  Members g*group .. g*group+group-1 are the repeats (seeds) of voltage vector g:
 *   sum[g,e]   = sum over the repeats of electrode_occ[m,e] / time[m]     (kmc_dopant_networks.py:618)
 *   sumsq[g,e] = sum of the squares,   count[g] = repeats with a finite time (count may be NULL)
 * so that mean and variance -- on one GPU or after an all-reduce of the three arrays -- need B/group*P values. */
int kmcb200_reduce_currents(int device, const double *time, const int64_t *electrode_occ, int64_t B, int P, int group,
                            double *sum /*[B/group,P]*/, double *sumsq /*[B/group,P]*/, double *count /*[B/group]*/, void *stream);

/* fp32 energies + dense rate matrix of ONE given state with the FAST kernel's arithmetic
 * (parity probe for the 1e-6-relative checks).  All host pointers.  site_energies_io[S]:
 * if energies_given != 0 the rates are evaluated AT these energies, otherwise the kernel's own
 * energies are computed from (E_constant, electrode_v, occupation) and written there. */
int kmcb200_probe_rates(kmcb200_layout *layout, const double *E_constant, const double *electrode_v,
                        double kT, const uint8_t *occupation, float *site_energies_io /*[S]*/,
                        int energies_given, float *rates_out /*[S*S]*/);

/* Micro-benchmarks of the pipes that bound the hop loop, measured on `device`:
 * what = 0 MUFU.EX2 (ex2/s), 1 FP32 FFMA (fma/s), 2 warp-instruction issue (warp-instructions/s),
 * 3 scattered 32-byte table lookups from L2 (sectors/s: the access pattern of the thread-per-trajectory kernel's hit path). */
double kmcb200_measure_peak(int device, int what);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
long long kmcb200_launch_count(void);

/* Name of the hop kernel the last kmcb200_run_ensemble of this process launched ("kmc_lanes_kernel",
 * "kmc_memo_kernel", "kmc_wide_kernel", "kmc_fast_kernel", "kmc_reforder_kernel", "kmc_exact_kernel",
 * "kmc_prob_kernel"; "" before the first call) -- bench.py names the kernel its roofline is about. */
const char *kmcb200_last_kernel(void);

#ifdef __cplusplus
}
#endif
#endif /* KMC_B200_H */
