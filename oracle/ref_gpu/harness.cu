// Test / bench infrastructure only: a harness around the reference's OWN CUDA solver, compiled from where it lies.
//
// `#include "implicit/gpu/als.cu"` below pulls in the unmodified /root/reference/implicit/gpu/als.cu
// (least_squares_cg_kernel :23-111, LeastSquaresSolver::calculate_yty :122-152, ::least_squares :154-197, with dot.cuh
// and convert.cuh) through -I /root/reference; nothing of it is copied into this repository.  The reference's
// Matrix / CSRMatrix wrappers (implicit/gpu/matrix.cu) need RMM, which is not installable here, so the include path
// also carries a 10-line stand-in for <rmm/device_uvector.hpp> (written by oracle/build_ref_gpu.py into its scratch
// directory) and this file supplies the two CSRMatrix members that als.cu's callers need, as plain pointer holders.
// Exposes one C entry point: a whole ALS-CG fit loop exactly as implicit/gpu/als.py:159-165 drives the solver
// (calculate_yty(Y) + least_squares(Cui, X, YtY, Y), then the same for the item side), timed with CUDA events.
#include "implicit/gpu/als.cu"

namespace implicit {
namespace gpu {
// stand-ins for implicit/gpu/matrix.cu:249-263 (no ownership: the harness owns the device arrays)
CSRMatrix::CSRMatrix(int rows, int cols, int nonzeros, const int *indptr_, const int *indices_, const float *data_)
    : indptr(const_cast<int *>(indptr_)), indices(const_cast<int *>(indices_)), data(const_cast<float *>(data_)), rows(rows),
      cols(cols), nonzeros(nonzeros) {}
CSRMatrix::~CSRMatrix() {}
// LeastSquaresSolver::calculate_loss (als.cu:253-281, not called by this harness) refers to these two members of
// implicit/gpu/matrix.cu; plain cudaMalloc stand-ins so that the shared object has no undefined symbol
Matrix::Matrix(size_t rows_, size_t cols_, void *host, bool allocate, size_t itemsize_)
    : rows(rows_), cols(cols_), data(host), itemsize(itemsize_) {
  if (allocate) {
    CHECK_CUDA(cudaMalloc(&data, rows * cols * itemsize));  // never freed: test infrastructure
    if (host) CHECK_CUDA(cudaMemcpy(data, host, rows * cols * itemsize, cudaMemcpyHostToDevice));
  }
}
void Matrix::to_host(void *output) const { CHECK_CUDA(cudaMemcpy(output, data, rows * cols * itemsize, cudaMemcpyDeviceToHost)); }
// implicit/gpu/als.cu declares the destructor in als.h and defines it near the end of the file
}  // namespace gpu
}  // namespace implicit

using implicit::gpu::CSRMatrix;
using implicit::gpu::LeastSquaresSolver;
using implicit::gpu::Matrix;

template <typename T>
static T *to_device(const T *host, size_t n) {
  T *d = nullptr;
  if (cudaMalloc(&d, sizeof(T) * (n ? n : 1)) != cudaSuccess) return nullptr;
  cudaMemcpy(d, host, sizeof(T) * n, cudaMemcpyHostToDevice);
  return d;
}

static Matrix view(float *dev, size_t rows, size_t cols) {
  Matrix m;  // the reference's default constructor: no storage, we point it at our device array
  m.rows = rows;
  m.cols = cols;
  m.data = dev;
  m.itemsize = 4;
  return m;
}

// returns 0 on success; ms_per_iteration[i] = device time of iteration i (both halves); X / Y are updated in place
extern "C" __attribute__((visibility("default"))) int ref_gpu_als_cg_fit(
    int users, int items, int factors, const int *ui_indptr, const int *ui_indices, const float *ui_data, int nnz,
    const int *iu_indptr, const int *iu_indices, const float *iu_data, float *X, float *Y, float regularization, int cg_steps,
    int iterations, float *ms_per_iteration) {
  try {
    int *d_uip = to_device(ui_indptr, (size_t)users + 1), *d_uix = to_device(ui_indices, nnz);
    float *d_uid = to_device(ui_data, nnz);
    int *d_iip = to_device(iu_indptr, (size_t)items + 1), *d_iix = to_device(iu_indices, nnz);
    float *d_iid = to_device(iu_data, nnz);
    float *d_X = to_device(X, (size_t)users * factors), *d_Y = to_device(Y, (size_t)items * factors), *d_G = nullptr;
    cudaMalloc(&d_G, sizeof(float) * factors * factors);
    if (!d_uip || !d_uix || !d_uid || !d_iip || !d_iix || !d_iid || !d_X || !d_Y || !d_G) return 2;
    {
      CSRMatrix Cui(users, items, nnz, d_uip, d_uix, d_uid), Ciu(items, users, nnz, d_iip, d_iix, d_iid);
      Matrix mX = view(d_X, users, factors), mY = view(d_Y, items, factors), mG = view(d_G, factors, factors);
      LeastSquaresSolver solver;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      for (int it = 0; it < iterations; ++it) {
        cudaEventRecord(e0);
        solver.calculate_yty(mY, &mG, regularization);           // implicit/gpu/als.py:160
        solver.least_squares(Cui, &mX, mG, mY, cg_steps);        // :161
        solver.calculate_yty(mX, &mG, regularization);           // :163
        solver.least_squares(Ciu, &mY, mG, mX, cg_steps);        // :164
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms_per_iteration[it], e0, e1);
      }
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    cudaMemcpy(X, d_X, sizeof(float) * (size_t)users * factors, cudaMemcpyDeviceToHost);
    cudaMemcpy(Y, d_Y, sizeof(float) * (size_t)items * factors, cudaMemcpyDeviceToHost);
    for (void *p : {(void *)d_uip, (void *)d_uix, (void *)d_uid, (void *)d_iip, (void *)d_iix, (void *)d_iid, (void *)d_X, (void *)d_Y, (void *)d_G})
      cudaFree(p);
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
  } catch (const std::exception &e) {
    fprintf(stderr, "ref_gpu harness: %s\n", e.what());
    return 1;
  }
}
