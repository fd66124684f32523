// kmc_internal.cuh -- shared declarations between the C-ABI host code and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kmcb200 {

// Device-resident tables of one layout (built once by kmcb200_layout_create).
struct LayoutDev {
    int N, P, S;
    int slots;      // ceil(S/32): row slots per lane in the fast kernel
    int pitch2;     // fast-table row pitch in float2 units (= 32*slots + 1, bank-conflict-free both ways)
    // fast table: tbl[j*pitch2 + i] = { nu*tc[i][j] (0 if pruned / diagonal / electrode-electrode),
    //                                   I_0*R/d[i][j] for acceptor-acceptor pairs i!=j, else 0 }
    // i.e. indexed [target j][source i]; fp32 narrowing as simulationWrapper.go:37-56.
    float2 *tbl;
    // production table (hop_fast.cu): acceptor sources only, tblf[j*pitchf + i], pitchf = 32*ceil(N/32)+1:
    //   target j <  N : { nu*tc[i][j], I_0*R/d[i][j] }      (i != j, else 0)
    //   target j >= N : { nu*tc[i][j], nu*tc[j][i] }        (acceptor i <-> electrode j-N, both directions)
    float2 *tblf;
    int pitchf;
    // sparse sweep of hop_wide.cu: near[j*32 + lane] = bits k: tblf[j*pitchf + lane + 32k] holds a non-zero constant (for an
    // electrode row: in either direction); sparse = 1 if at most a third of the acceptor-acceptor constants are non-zero
    unsigned char *near;
    int sparse;
    // replay tables (row-major, exactly the caller's values)
    float *d32, *tc32;       // [S*S] narrowed (Go semantics)
    double *d64, *tc64;      // [S*S] (numba semantics)
    int2 *pairs;             // Go transition list (simulation.go:199-215), length L
    int L;
    float nu32, I032, R32;   // narrowed scalars (simulationWrapper.go:92)
    double nu64, I064, R64;
};

struct EnsembleDev {
    int64_t B, hops, prehops;
    int mode;
    const double *E_constant, *basis, *electrode_v, *kT;
    const uint8_t *occupation0;
    uint64_t seed, member_index0;
    const double *stream_e; const float *stream_u; const double *stream_u64;
    double *time; int64_t *electrode_occ; uint8_t *occupation_out; double *site_energies_out;
    double *avg_occupation; double *traffic; int32_t *trace;
    long long *misses;  // [B] rate-structure evaluations (cache misses) per member, or null
    double *prob_occupation, *prob_electrode_occ;  // MODE_PROB: fractional occupations [B,N], electrode tallies [B,P]
    double *scratch;   // replay kernels: [B][S*S] doubles (rate / cumulative list)
    unsigned long long *queue;  // memoised kernel: member work queue (zeroed before the launch)
    unsigned char *gtab;  // memoised kernels: second-level cache, warp_slots * 2^gtab_log entries of 288 / 448 B (or null);
                          // hop_lanes.cu: the state table, warp_slots * 2^gtab_log sets of 128 B
    int gtab_log;         // log2(entries / sets per warp slot); 0 = no second level
    int lanes_flags;      // hop_lanes.cu: bit 0 = do not consult the table (every hop is evaluated; for testing)
    uint32_t launch_id;   // hop_memo.cu / hop_wide.cu: entries are valid only with the tag (launch_id, member + 1): the table
                          // is zeroed ONCE, at allocation, and never reset -- neither per launch nor per member
    // hop_lanes.cu: Philox round keys (key + r * Weyl constants, r = 0..9) and the slicing of the queue's tail: blocks
    // [0, lanes_nb_full) run all their hops as one work item, the others in lanes_ns slices of lanes_slice_hops hops
    // (a multiple of 64), handed out slice-major; lanes_prog[block - lanes_nb_full] = slices done, lanes_ck[member] =
    // occupation mask between slices (time and tallies are checkpointed in the output arrays)
    uint32_t rk[20];
    int64_t lanes_nb_full, lanes_slice_hops;
    int lanes_ns;
    int solo_emax;       // kmc_solo_kernel: entries of a trajectory's state graph in shared memory (set by launch_solo)
    int lanes_halves;    // 1: every block of 32 members is handed out twice and split where a run starts at member 16 (hop_lanes.cu)
    uint32_t *lanes_prog, *lanes_ck;
};

// launchers (return cudaError_t of the launch)
cudaError_t launch_fast(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches);
struct MemoPlan { int64_t warp_slots; int64_t max_slots; };  // (max_slots: hop_lanes.cu, warp slots of a full device)
cudaError_t launch_memo(const LayoutDev &L, const EnsembleDev &E, int logk, cudaStream_t st, int *launches,
                        MemoPlan *plan_only = nullptr);
cudaError_t launch_wide(const LayoutDev &L, const EnsembleDev &E, int logk, cudaStream_t st, int *launches,
                        MemoPlan *plan_only = nullptr);
cudaError_t launch_solo(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches);
cudaError_t launch_lanes(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches,
                         MemoPlan *plan_only = nullptr);
cudaError_t launch_reforder(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches);
cudaError_t launch_prob(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches);
cudaError_t launch_exact(const LayoutDev &L, const EnsembleDev &E, cudaStream_t st, int *launches);
cudaError_t launch_probe(const LayoutDev &L, const double *E_constant, const double *electrode_v, double kT,
                         const uint8_t *occ, float *se_io, int se_given, float *rates_out, cudaStream_t st,
                         int *launches);

double measure_peak(int what, int *launches);
cudaError_t launch_reduce_currents(const double *time, const int64_t *eo, int64_t B, int P, int group, double *sum, double *sumsq,
                                   double *count, cudaStream_t st, int *launches);

}  // namespace kmcb200
