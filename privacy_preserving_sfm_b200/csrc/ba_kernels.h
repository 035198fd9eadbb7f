// ba_kernels.h — device-side data model and launch wrappers of the bundle-adjustment kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace ppsfm {

// All arrays live in HBM.  K = kept observations (point-major), C = images, P = points,
// NB = camera blocks (images with a variable pose that have observations).
struct BaDev {
  int C = 0, P = 0, NB = 0, n = 0, ld = 0;
  int has_ext_models = 0;  // a camera of model id >= 5 is present (out-of-line model evaluation)
  int64_t K = 0;
  // static structure
  int* obs_cam = nullptr;       // [K] image index
  int* obs_pt = nullptr;        // [K] point index
  double* obs_line = nullptr;   // [3][K] SoA
  int64_t* pt_start = nullptr;  // [P+1] first observation of every point
  int* cam_obs = nullptr;       // [K] observation ids grouped by camera block
  int64_t* cam_start = nullptr; // [NB+1]
  int* block_img = nullptr;     // [NB] image index of a block
  int* cam_block = nullptr;     // [C] block index or -1 (constant pose)
  uint8_t* cam_mask = nullptr;  // [C] bit a set = tangent dim a (rot 0..2, trans 3..5) is free
  int* img_model = nullptr;     // [C] COLMAP camera model id
  double* img_params = nullptr; // [C][12] intrinsics (constant), expanded per image
  int num_cameras = 0;
  int* img_cam = nullptr;       // [C] camera index of an image
  int* cam_model = nullptr;     // [num_cameras]
  double* cam_params = nullptr; // [num_cameras][12]
  uint8_t* pt_var = nullptr;    // [P]
  // intrinsics refinement (ba_intrinsics.cu); NCv = 0: every camera constant, nothing below is used
  int NCv = 0;                      // cameras with at least one variable parameter
  int* cam_intr_block = nullptr;    // [num_cameras] intrinsics block of a camera or -1
  unsigned* intr_mask = nullptr;    // [NCv] bit a = parameter a is variable
  int* cam_nparams = nullptr;       // [num_cameras] CameraModel::kNumParams
  double* intr_scale = nullptr;     // [NCv][12] Jacobi scales
  double* Ji = nullptr;             // [K][2][12] d r / d params (scaled, loss-corrected, masked)
  double* Uii = nullptr;            // [NCv][12][12]   (Uii | Uic | gi are contiguous behind U | gc)
  double* Uic = nullptr;            // [NB][12][6] coupling with the pose block of the image
  double* gi = nullptr;             // [NCv][12]
  // state
  double* q = nullptr;   // [C][4]
  double* t = nullptr;   // [C][3]
  double* X = nullptr;   // [P][3]
  double* qn = nullptr;  // candidate
  double* tn = nullptr;
  double* Xn = nullptr;
  // linearisation (scaled by the Jacobi column scales, loss-corrected): 20 fields per
  // observation — residual r (2), J_c row-major 2x6 [rot | trans] (12), J_p row-major 2x3 (6) —
  // stored "blocked SoA": [chunk of 256 observations][field][256].  Every warp access is still a
  // coalesced 256-byte line, but one chunk is a single contiguous 40 KB region instead of 20
  // streams 16 MB apart (DRAM row-buffer locality for the write-heavy Jacobian build).
  double* J = nullptr;    // [ceil(K/256)][20][256]
  double* cam_scale = nullptr;  // [NB][6]
  double* pt_scale = nullptr;   // [P][3]
  // normal equations
  double* U = nullptr;    // [NB][36]
  double* Upart = nullptr; // [NB][4][27] partial sums of the camera normal equations
  double* gc = nullptr;   // [NB][6]
  double* V = nullptr;    // [6][P] symmetric (00 01 02 11 12 22), SoA
  double* gp = nullptr;   // [3][P]
  double* Vinv = nullptr; // [6][P]
  double* Lv = nullptr;   // [P][9] inverse Cholesky factor of the damped V (6) and h = L^-1 g_p (3)
  double* Zrec = nullptr; // [K][24] per observation: Z = J_c^T J_p L^-T (6x3 row-major), z = Z h (6)
  // Schur gather lists (static structure, ba_schur.cu): observation pairs grouped by camera pair
  int sch_npairs = 0, sch_nchunks = 0, sch_nmulti = 0;
  int2* sch_ent = nullptr;          // [entries] (e, f), sorted by camera pair (i >= j), then e
  int64_t* sch_pair_start = nullptr;  // [npairs + 1]
  int* sch_pair_chunk = nullptr;    // [npairs + 1] first chunk (<= 128 entries) of a pair
  int* sch_chunk_pair = nullptr;    // [nchunks]
  int* sch_multi = nullptr;         // [nmulti] pairs that span several chunks
  double* sch_partial = nullptr;    // [nchunks][48] chunk sums of those pairs
  double* S = nullptr;    // [ld][ld] bordered reduced camera matrix (row n = rhs)
  double* Spacked = nullptr;  // multi-GPU: packed lower triangle + rhs for the all-reduce
  double* dc = nullptr;   // [n]
  double* dp = nullptr;   // [3][P]
  double* u = nullptr;    // [2][K] J_c dc per observation (back-substitution scratch)
  int* chol_status = nullptr;
  double* chol_work = nullptr;  // 64x64 scratch of the dense solver
  // reductions
  double* partials = nullptr;  // scratch for block partial sums
  int num_partials = 0;
  double* scalars = nullptr;   // [16] device scalars
};

// Device copies of the caller's arrays (input order), consumed by the assembly (ba_assembly.cu).
struct BaRaw {
  int C = 0, P = 0;  // P: points of the WHOLE problem (the caller's indices)
  // sharded solve: this rank owns the points p with p % world == rank and numbers them p / world;
  // every per-point array on the device has P_local entries (world == 1: P_local == P)
  int P_local = 0, rank = 0, world = 1;
  int64_t O = 0;
  const int* obs_image = nullptr;       // [O]
  const int* obs_point = nullptr;       // [O]
  const double* obs_line = nullptr;     // [O][3]
  const uint8_t* pose_flags = nullptr;  // [C] (zeros if the caller passed none)
  const uint8_t* point_const = nullptr; // [P]
};
cudaError_t ba_assemble_points(BaDev& d, const BaRaw& raw, int rank, int world,
                               void* (*alloc)(void*, size_t), void* alloc_ctx, cudaStream_t s,
                               uint8_t* cam_used_host, int64_t* first_bad_index,
                               int64_t* first_bad_norm);
cudaError_t ba_assemble_cameras(BaDev& d, void* (*alloc)(void*, size_t), void* alloc_ctx,
                                cudaStream_t s);

// reduced-system width of one intrinsics block (the widest models have 12 parameters)
constexpr int kIntrW = 12;
constexpr int kMaxVarCams = 64;
inline size_t intr_normal_doubles(int NB, int NCv) {
  return (size_t)NCv * kIntrW * kIntrW + (size_t)NB * kIntrW * 6 + (size_t)NCv * kIntrW;
}

constexpr int kJFields = 20;
__host__ __device__ inline size_t ba_jidx(int field, int64_t k) {
  return ((size_t)(k >> 8) * kJFields + (size_t)field) * 256 + (size_t)(k & 255);
}
inline size_t ba_j_doubles(int64_t K) { return (size_t)((K + 255) >> 8) * kJFields * 256; }

enum BaScalar { kCost = 0, kModelChange = 1, kStepSq = 2, kXSq = 3, kGradMax = 4, kNumScalars = 16 };

struct BaLoss {
  int type;      // 0 TRIVIAL, 1 SOFT_L1, 2 CAUCHY
  double scale;
};

// residuals (+ Jacobians) at (q, t, X) given as pointers so that the candidate state can be
// evaluated; cost = 0.5 sum rho(|r|^2) is left in scalars[kCost].  Returns #launches.
int launch_linearize(const BaDev& d, const double* q, const double* t, const double* X,
                     bool jacobians, BaLoss loss, cudaStream_t s);
// U, gc, V, gp from the stored linearisation.
int launch_normal_equations(const BaDev& d, cudaStream_t s);
// Jacobi scales 1 / (1 + sqrt(diag)) from U, V (computed with unit scales).
int launch_jacobi_scales(const BaDev& d, cudaStream_t s);
// One-time structure set-up of the reduced-system assembly (gather lists, record storage);
// alloc(ctx, bytes) returns device memory owned by the problem (nullptr on failure).
cudaError_t build_schur_lists(BaDev& d, void* (*alloc)(void*, size_t), void* alloc_ctx,
                              cudaStream_t s);
// S = blockdiag(U + D_c^2) - W (V + D_p^2)^-1 W^T (lower triangle), row n = rhs; stores Vinv.
// Every block of the lower triangle and the rhs row are overwritten (no memset needed).
// `include_camera_terms` is false on ranks > 0 of a sharded solve (the replicated camera terms
// must enter the all-reduced sum once).
int launch_build_reduced_system(const BaDev& d, double radius, double min_diag, double max_diag,
                                bool include_camera_terms, cudaStream_t s);
// dp from dc, model cost change, candidate state, step / x norms (scalars).
int launch_backsubstitute_and_update(const BaDev& d, bool count_camera_norms, cudaStream_t s);
// lower triangle + rhs row of the bordered reduced matrix <-> packed buffer (multi-GPU all-reduce)
size_t packed_lower_doubles(int n);
void launch_pack_lower(const double* S, int n, int ld, double* packed, cudaStream_t s);
void launch_unpack_lower(double* S, int n, int ld, const double* packed, cudaStream_t s);
// out = a + beta * b
void launch_scatter_displacement(double* Xg, const double* X, const double* X0, int P, int world,
                                 int rank, cudaStream_t s);
void launch_axpby(double* out, const double* a, const double* b, double beta, size_t n,
                  cudaStream_t s);
// max |x - Plus(x, -g)| over all blocks -> scalars[kGradMax]
int launch_gradient_max_norm(const BaDev& d, cudaStream_t s);

// Intrinsics refinement (ba_intrinsics.cu); every launcher is a no-op returning 0 when d.NCv == 0.
int launch_intr_jacobian(const BaDev& d, const double* q, const double* t, const double* X,
                         BaLoss loss, cudaStream_t s);
int launch_intr_normal(const BaDev& d, cudaStream_t s);           // Uii, Uic, gi (zeroed first)
int launch_intr_scales(const BaDev& d, bool jacobi, cudaStream_t s);
// intrinsics rows of S (after launch_build_reduced_system); *overflow = 1 if a point sees more
// distinct variable cameras than the kernel holds
int launch_intr_reduced_rows(const BaDev& d, double radius, double min_diag, double max_diag,
                             bool include_camera_terms, int* overflow, cudaStream_t s);
int launch_intr_backsub(const BaDev& d, cudaStream_t s);          // u += J_i d_i, acc_p += J_p^T (J_i d_i)
// candidate parameters (per camera and expanded per image); adds to scalars[kStepSq / kXSq]
int launch_intr_update(const BaDev& d, double* cam_params_n, double* img_params_n,
                       bool count_norms, cudaStream_t s);
int launch_intr_gradient(const BaDev& d, cudaStream_t s);

}  // namespace ppsfm
