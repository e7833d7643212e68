/*
 * aceb200.h -- C ABI of the B200-native ACE evaluation hot path.
 *
 * This is the drop-in boundary under ACE.jl's evaluation methods.  ACE.jl has
 * no FFI of its own; its extension seam is multiple dispatch on the `evaluator` field of
 * `LinearACEModel` (src/linearmodel.jl:36-40, 107-111, 133-134) and on the basis types.  Each entry
 * point below names the reference method(s) it replaces.  Host code (Julia `ccall`, or the ctypes
 * mirror in ace_jl_b200/) fills `aceb200_desc` from the live basis/model objects once; all tables
 * are copied to the GPU at `aceb200_model_create`.
 *
 * Conventions
 *   - every function returns 0 on success or a negative ACEB200_E* code; the message is available
 *     from aceb200_last_error() (thread-local).  Nothing aborts.
 *   - all index tables are 1-based exactly as the Julia objects hold them (converted once, on
 *     upload); integers are int32 unless stated.
 *   - complex numbers are interleaved (re, im) doubles.
 *   - the caller owns every buffer and keeps it alive for the duration of the call; no pointer is
 *     retained after a call returns (tables are copied at create).
 *   - a model handle is immutable except through aceb200_set_params(); evaluation calls on one
 *     handle may be issued from several host threads (they serialise on the handle's workspace).
 *   - there is NO CPU fallback: without a CUDA device every call fails with ACEB200_ECUDA.
 */
#ifndef ACEB200_H
#define ACEB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACEB200_ABI_VERSION 2

/* error codes */
#define ACEB200_OK            0
#define ACEB200_EDESC        -1   /* malformed descriptor / batch                         */
#define ACEB200_EUNSUPPORTED -2   /* transform / component / size outside the supported set */
#define ACEB200_ECUDA        -3   /* CUDA runtime error (including "no device")            */
#define ACEB200_ENOMEM       -4   /* out of device or host memory                          */
#define ACEB200_EEMPTY       -5   /* an environment with zero neighbours where the reference
                                     asserts length(cfg) > 0 (src/product_1pbasis.jl:124)     */
#define ACEB200_ECATEGORY    -6   /* species code outside 1..n_cat (src/discrete1pbasis.jl:39) */

/* distance transforms: the closed set of src/transforms/distancetransforms.jl:16-25 */
#define ACEB200_TRANS_ID     0    /* t = r                                  par: -                */
#define ACEB200_TRANS_POLY   1    /* t = ((1+r0)/(1+r))^p                   par: p, r0            */
#define ACEB200_TRANS_MORSE  2    /* t = exp(-lambda (r/r0 - 1))            par: lambda, r0       */
#define ACEB200_TRANS_AGNESI 3    /* t = 1/(1 + a (r/r0)^p)                 par: r0, p, a         */

/* one-particle basis component kinds, in the order of Product1pBasis.bases (src/product_1pbasis.jl:5-8) */
#define ACEB200_COMP_RN   0       /* Rn1pBasis   (src/b1pcomponents/Rn.jl:17-29)   */
#define ACEB200_COMP_YLM  1       /* Ylm1pBasis  (src/b1pcomponents/Ylm.jl:18-26)  */
#define ACEB200_COMP_CAT  2       /* Categorical1pBasis (src/discrete1pbasis.jl:58) */
#define ACEB200_MAX_COMP  4

/* memory space of batch inputs and of the outputs of a call */
#define ACEB200_HOST   0
#define ACEB200_DEVICE 1

typedef struct aceb200_model aceb200_model;

/* Everything a LinearACEModel / SymmetricBasis holds that evaluation needs. */
typedef struct aceb200_desc {
    int32_t abi_version;        /* = ACEB200_ABI_VERSION */
    int32_t struct_bytes;       /* = sizeof(aceb200_desc) */

    /* radial basis: OrthPolyBasis (src/polynomials/orthpolys.jl:82-92) of the transformed distance */
    int32_t n_rad;              /* length(J) = N                                  */
    int32_t pl, pr;             /* envelope powers                                */
    double  tl, tr;             /* envelope roots (transformed variable)          */
    const double *rad_A;        /* [n_rad] recursion coefficients                 */
    const double *rad_B;        /* [n_rad]                                        */
    const double *rad_C;        /* [n_rad]                                        */
    int32_t trans_kind;         /* ACEB200_TRANS_*, parsed from Lambda.exstr (src/transforms/lambdas.jl:9-12) */
    int32_t _pad0;
    double  trans_par[4];

    /* angular basis: SHBasis(maxL) (src/polynomials/sphericalharmonics.jl:285-291) */
    int32_t maxL;
    /* categorical basis: number of categories Q, 0 if there is no Categorical1pBasis */
    int32_t n_cat;

    /* Product1pBasis (src/product_1pbasis.jl:5-8) */
    int32_t n_comp;                         /* NB                                          */
    int32_t comp_kind[ACEB200_MAX_COMP];    /* kind of bases[i]                            */
    int32_t nA;                             /* length(basis1p)                             */
    const int32_t *indices;                 /* [nA][n_comp] = Vector{NTuple{NB,Int}}, 1-based */

    /* PIBasisSpec (src/pibasis.jl:10-13) */
    int32_t nAA;
    int32_t maxord;                         /* size(iAA2iA, 2)                             */
    const int32_t *orders;                  /* [nAA]                                       */
    const int32_t *iAA2iA;                  /* column-major nAA x maxord (Julia Matrix), 1-based, 0 = padding */
    int32_t pireal;                         /* PIBasis.real === real (src/pibasis.jl:148)  */
    int32_t symreal;                        /* SymmetricBasis.real === real (src/symmbasis.jl:37) */

    /* A2Bmap::SparseMatrixCSC{PROP,Int} (src/symmbasis.jl:33-38), nB x nAA */
    int32_t nB;
    int32_t ncomp;                          /* complex components per entry: 1, 3 or 9     */
    int64_t nnz;
    const int32_t *colptr;                  /* [nAA+1], 1-based                            */
    const int32_t *rowval;                  /* [nnz],   1-based                            */
    const double  *nzval;                   /* [nnz][ncomp] complex                        */

    /* LinearACEModel.c (src/linearmodel.jl:36-40): Vector{T} or Vector{SVector{nprop,T}} */
    int32_t nprop;
    int32_t _pad1;
    const double *c;                        /* [nB][nprop] real; may be NULL (zeros)       */
} aceb200_desc;

/* A ragged batch of atomic environments.  R has exactly the memory of
 * Vector{PositionState{Float64}} (src/states.jl:396: isbits, 24-byte stride) for the
 * concatenated configurations. */
typedef struct aceb200_batch {
    int64_t nenv;
    const int64_t *offsets;     /* [nenv+1], offsets[0] = 0; environment e owns neighbours offsets[e]..offsets[e+1]-1 */
    const double  *R;           /* [offsets[nenv]][3]                                                              */
    const int32_t *species;     /* [offsets[nenv]] 1-based category index (val2i, src/discrete1pbasis.jl:33), or NULL */
    int32_t space;              /* ACEB200_HOST or ACEB200_DEVICE: where offsets/R/species AND the outputs live     */
    int32_t _pad;
    int64_t nJ;                 /* total neighbour count offsets[nenv] (= length of R) if the caller knows it, else 0.  A
                                   DEVICE batch with nJ > 0 is evaluated without any device -> host round trip before
                                   the kernels are launched; with nJ = 0 the library reads offsets[nenv] back first.       */
} aceb200_batch;

/* ---- lifecycle -------------------------------------------------------------------------- */

/* number of CUDA devices visible; <= 0 means the library cannot run */
int aceb200_device_count(void);
/* device used by models created afterwards from this host thread (default 0) */
int aceb200_set_device(int device);
/* Evaluate HOST batches of this model on several GPUs: `devices` lists n distinct CUDA devices and must contain the one
 * the model was created on.  The tables are replicated once; every later HOST-batch call cuts the batch into n contiguous
 * shards balanced by neighbour count, evaluates them concurrently (one host thread, stream set and copy pipeline per
 * device) and writes each shard's slice of the caller's buffers.  DEVICE batches and structures stay on the model's own
 * device (their memory lives there).  n = 1 with the model's device drops the replicas.  (in the
 * reference the same effect needs `Threads.@threads` over configurations, src/utils/pools.jl:44-75.) */
int aceb200_set_devices(aceb200_model *m, int n, const int *devices);
/* copy `n` bytes of the calling thread's last error message */
int aceb200_last_error(char *buf, int n);

/* Replaces the construction of a ProductEvaluator (src/evaluator.jl:30-31) plus the read-only use of
 * basis1p / pibasis.spec / A2Bmap by every evaluate method.  Uploads all tables, computes c~. */
int aceb200_model_create(const aceb200_desc *desc, aceb200_model **out);
int aceb200_model_destroy(aceb200_model *m);
/* set_params!(m, c) -> _get_eff_coeffs! (src/linearmodel.jl:67-71, src/evaluator.jl:48-66); c is [nB][nprop] host */
int aceb200_set_params(aceb200_model *m, const double *c, int64_t n);
/* read back c~ = A2Bmap^T c as [nAA][nprop][ncomp] complex (ProductEvaluator.coeffs, src/evaluator.jl:11) */
int aceb200_get_eff_coeffs(aceb200_model *m, double *ctilde);
/* run subsequent DEVICE-resident calls on this CUDA stream (a cudaStream_t); NULL = the legacy default
 * stream.  HOST-resident batches are pipelined over three private streams and return when all copies are done. */
int aceb200_set_stream(aceb200_model *m, void *cuda_stream);
/* number of kernels this handle has launched so far (monotone counter, for audits) */
int64_t aceb200_launch_count(const aceb200_model *m);

/* ---- basis values ------------------------------------------------------------------------ */

/* evaluate(basis1p::Product1pBasis, cfg) (src/product_1pbasis.jl:123-134): A [nenv][nA] complex */
int aceb200_eval_A(aceb200_model *m, const aceb200_batch *b, double *A);
/* evaluate(pibasis::PIBasis, cfg) (src/pibasis.jl:258-294): AA [nenv][nAA], real if pireal else complex */
int aceb200_eval_AA(aceb200_model *m, const aceb200_batch *b, double *AA);
/* evaluate(basis::SymmetricBasis, cfg) (src/symmbasis.jl:297-316): B [nenv][nB][ncomp], real if symreal else complex */
int aceb200_eval_B(aceb200_model *m, const aceb200_batch *b, double *B);

/* ---- basis Jacobians (evaluate_d / evaluate_ed) -------------------------------------------- */
/* Per environment these are Julia Matrix{DState}(nbasis x J), column-major: [neighbour][basis index][...]. */

/* evaluate_ed(basis1p, cfg) (src/product_1pbasis.jl:234-250): A as above (may be NULL),
 * dA [sum J][nA][3] complex */
int aceb200_eval_dA(aceb200_model *m, const aceb200_batch *b, double *A, double *dA);
/* evaluate_ed(pibasis, cfg) (src/pibasis.jl:303-332, 402-432): dAA [sum J][nAA][3], real if pireal else complex */
int aceb200_eval_dAA(aceb200_model *m, const aceb200_batch *b, double *AA, double *dAA);
/* evaluate_ed(basis, cfg) (src/symmbasis.jl:322-336): dB [sum J][nB][3][ncomp] (component fastest,
 * the memory of coco_o_daa's result, src/properties.jl:53-59), real if symreal else complex */
int aceb200_eval_dB(aceb200_model *m, const aceb200_batch *b, double *B, double *dB);

/* ---- linear model ------------------------------------------------------------------------- */

/* evaluate(m::LinearACEModel, cfg) via ProductEvaluator (src/evaluator.jl:121-147):
 * E [nenv][nprop][ncomp], real if symreal else complex */
int aceb200_energy(aceb200_model *m, const aceb200_batch *b, double *E);
/* evaluate + grad_config (src/evaluator.jl:150-200): G [sum J][nprop][3][ncomp] (component fastest),
 * real if symreal else complex; E may be NULL */
int aceb200_energy_forces(aceb200_model *m, const aceb200_batch *b, double *E, double *G);
/* _rrule_evaluate(dp::SVector{nprop}, m, V, cfg) (src/evaluator.jl:161-200; `contract(dp, c~[iAA])` at :183; the
 * benchmark's call, benchmark/bm_linear.jl:92-93): the pullback of a multi-property model contracted over the
 * properties,  G [sum J][3][ncomp] = sum_p dp[p] * (gradient of property p).  The adjoint pass runs over all properties,
 * its result is contracted on the device, and ONE force field is assembled (for 16 properties: 1/16 of the force work and
 * of the output bytes of aceb200_energy_forces).  dp: [nprop] HOST doubles (nprop <= 32); E [nenv][nprop][ncomp] may be NULL. */
int aceb200_energy_forces_dp(aceb200_model *m, const aceb200_batch *b, const double *dp, double *E, double *G);
/* grad_params(m, cfg) (src/linearmodel.jl:114-123) is eval_B; grad_params_config (:127) is eval_dB. */
/* adjoint_EVAL_D(m, cfg, w) (src/evaluator.jl:204-244; src/linearmodel.jl:133-134):
 * out_k = A2Bmap * real?( sum_t (sum_j w_j . grad phi_{v_t}(r_j)) prod_{s != t} A_{v_s} ), i.e. sum_j w_j . dB_k/dr_j
 * without forming a Jacobian.  w: [sum J][3] real (DState rr per neighbour), in the batch's memory space;
 * out: [nenv][nB][ncomp] complex. */
int aceb200_adjoint_eval_d(aceb200_model *m, const aceb200_batch *b, const double *w, double *out);

/* ---- caller side: a whole atomic structure ------------------------------ */

/* An atomic structure with its neighbour list, sorted by centre.  Replaces, around the per-environment
 * calls above, the loop that JuLIP / ACEatoms.jl run on the host (energy(V, at), forces(V, at), virial(V, at):
 * for each centre i, Rs = {x_j + S_ij - x_i}, dV = evaluate_d(V, Rs), frc[j] -= dV_j, frc[i] += dV_j,
 * vir -= dV_j (x) R_j).  Environments are built on the device from X and the pair table, and the pair
 * gradients never leave it: per pair the interface moves 4 B (+ 3 B with periodic images, + 4 B with a
 * reverse table) instead of the 48 B of aceb200_energy_forces.  The neighbour list must be a FULL list (every
 * pair appears under both of its centres), as JuLIP's is. */
#define ACEB200_NBR_PACKED 1

typedef struct aceb200_structure {
    int64_t natoms;
    int64_t npairs;           /* < 2^31 */
    const double  *X;         /* [natoms][3] positions                                                          */
    const int64_t *first;     /* [natoms+1] pairs first[i]..first[i+1]-1 have centre i; first[0] = 0              */
    const int32_t *nbr;       /* [npairs] 0-based index j of the neighbour atom                                   */
    const int8_t  *image;     /* [npairs][3] integer image shift S (JuLIP's neighbour list: i, j, S):
                                 R = X[j] + S[0] cell[0] + S[1] cell[1] + S[2] cell[2] - X[i]; NULL = no periodic images */
    const int32_t *species;   /* [natoms] 1-based category of each atom (a pair takes its neighbour's), or NULL   */
    const int32_t *rev;       /* [npairs] index of the reverse pair (centre j, neighbour i, image -S), -1 if there is
                                 none (the neighbour is not a centre); NULL = found on the device by searching j's pairs */
    double cell[9];           /* lattice vectors as rows, HOST memory in either space; unused when image is NULL   */
    int32_t space;            /* ACEB200_HOST or ACEB200_DEVICE: where every pointer above AND the outputs live   */
    int32_t flags;            /* 0, or ACEB200_NBR_PACKED: nbr[p] = j | (S0+1) << 26 | (S1+1) << 28 | (S2+1) << 30 with image = NULL
                                 (natoms <= 2^26, S in {-1, 0, 1}: any cell wider than the cutoff): 4 instead of 7 bytes per
                                 pair cross PCIe, which is what bounds this call when several GPUs share one host          */
} aceb200_structure;

/* Site energies, atomic forces and the virial of a structure for a real (symreal) model:
 *   Esite [natoms][nprop][ncomp]
 *   F     [natoms][nprop][3][ncomp]   F_i = -dE/dx_i = sum_{p in env(i)} (g_p - g_rev(p)): a gather, no atomics,
 *                                      bit-reproducible
 *   W     [nprop][3][3] or NULL       W_ab = -sum_p g_p[a] R_p[b]   (needs ncomp = 1)
 * Esite and W may be NULL. */
int aceb200_structure_energy_forces(aceb200_model *m, const aceb200_structure *s, double *Esite, double *F, double *W);

/* ---- introspection (sizes the host needs to allocate outputs) ------------------------------ */
typedef struct aceb200_sizes {
    int32_t nA, nAA, nB, ncomp, nprop, maxord, pireal, symreal;
} aceb200_sizes;
int aceb200_model_sizes(const aceb200_model *m, aceb200_sizes *out);

/* Device-side timing of the last evaluation call on this handle, in milliseconds: CUDA events
 * recorded on the handle's stream around the kernel launches only (no copies). */
int aceb200_last_kernel_ms(const aceb200_model *m, double *ms);

/* per-kernel split of the last aceb200_energy / aceb200_energy_forces call: ms3 = {pool, adjoint, forces} */
int aceb200_last_stage_ms(const aceb200_model *m, double *ms3);

/* Measured FP64 FMA throughput of the current device in TFLOP/s (a register-resident FMA probe).
 * The FP64 pipe is the roofline that bounds this path and the driver's
 * MEASURED_PEAKS.json does not record it. */
int aceb200_measure_fp64(double *tflops);

/* Measured FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) throughput of the current device in TFLOP/s: the
 * roofline the north_star names for the dense multi-property readout (src/evaluator.jl:137-143 with SVector
 * coefficients); reported next to the FMA figure so that the choice between the two pipes is made on data. */
int aceb200_measure_dmma(double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* ACEB200_H */
