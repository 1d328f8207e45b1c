// XLA-FFI handlers over the C-ABI (include/tn4ml_b200.h): the symbols `jax.ffi.register_ffi_target` binds when tn4ml's
// JAX loop calls this library.  Replaces loss_fn / value_and_grad of
// tn4ml/models/model.py:723-751, :554-566 as a jax.custom_vjp pair.
//
// Compiled only where the XLA FFI headers exist (jaxlib ships them under jaxlib/include; add that directory with -I).
// This image has neither jax nor the headers, so the release library exports tnb_ffi_available() == 0 and the handlers
// below are NOT built or tested here; everything they forward to is exercised through ctypes (tests/test_gpu_*.py).
#include <stdint.h>

#include "../../include/tn4ml_b200.h"

#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define TNB_HAVE_XLA_FFI 1
#endif
#endif

#ifndef TNB_HAVE_XLA_FFI

extern "C" int32_t tnb_ffi_available(void) { return 0; }

#else  // ---------------------------------------------------------------------------------------------------------

#include <cuda_runtime.h>

#include <cstring>
#include <string_view>

#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

extern "C" int32_t tnb_ffi_available(void) { return 1; }

// The static description travels as one opaque attribute: the bytes of a tnb_problem (what tn4ml_b200._lib.TnbProblem
// serialises; the Python side builds it with ctypes, so both ends share the header's layout).
static ffi::Error problem_of(std::string_view bytes, tnb_problem *out) {
  if (bytes.size() != sizeof(tnb_problem))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "attribute `problem` is not a tnb_problem");
  std::memcpy(out, bytes.data(), sizeof(tnb_problem));
  return ffi::Error::Success();
}

static ffi::Error status(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, tnb_last_error());
}

static int32_t dtype_of(ffi::DataType t) { return t == ffi::DataType::F64 ? TNB_F64 : TNB_F32; }

// tnb_forward: x [B, L], targets [B, C] (or a zero-sized buffer), packed params ->
//   loss [B], loss_sum f64[1], out [B, C], out_exp i32[B], workspace u8[tnb_workspace_bytes(prob, B, 1)]
// The workspace result is the VJP residual (jax.custom_vjp passes it back to tnb_backward_ffi).
static ffi::Error ForwardImpl(cudaStream_t stream, ffi::AnyBuffer x, ffi::AnyBuffer targets, ffi::AnyBuffer params,
                              ffi::Result<ffi::AnyBuffer> loss, ffi::Result<ffi::AnyBuffer> loss_sum,
                              ffi::Result<ffi::AnyBuffer> out, ffi::Result<ffi::AnyBuffer> out_exp,
                              ffi::Result<ffi::AnyBuffer> workspace, std::string_view problem) {
  tnb_problem p;
  if (ffi::Error e = problem_of(problem, &p); e.failure()) return e;
  if (x.dimensions().size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "x must be [B, L]");
  if (dtype_of(x.element_type()) != p.dtype) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "dtype mismatch");
  const int64_t B = x.dimensions()[0];
  const void *t = targets.element_count() > 0 ? targets.untyped_data() : nullptr;
  if (cudaMemsetAsync(loss_sum->untyped_data(), 0, sizeof(double), stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemsetAsync(loss_sum) failed");
  return status(tnb_forward(&p, B, x.untyped_data(), t, params.untyped_data(), loss->untyped_data(),
                            static_cast<double *>(loss_sum->untyped_data()), out->untyped_data(),
                            static_cast<int32_t *>(out_exp->untyped_data()), /*keep_stash=*/1, workspace->untyped_data(),
                            static_cast<int64_t>(workspace->size_bytes()), stream));
}

// tnb_backward: x, targets, packed params, workspace (residual of the forward call) -> packed grads of
// loss_scale * sum_b loss_b
static ffi::Error BackwardImpl(cudaStream_t stream, ffi::AnyBuffer x, ffi::AnyBuffer targets, ffi::AnyBuffer params,
                               ffi::AnyBuffer workspace, ffi::Result<ffi::AnyBuffer> grads, std::string_view problem,
                               double loss_scale) {
  tnb_problem p;
  if (ffi::Error e = problem_of(problem, &p); e.failure()) return e;
  const int64_t B = x.dimensions()[0];
  const void *t = targets.element_count() > 0 ? targets.untyped_data() : nullptr;
  // the library only reads and rewrites scratch regions of the workspace; XLA hands the residual over as an input
  return status(tnb_backward(&p, B, x.untyped_data(), t, params.untyped_data(), loss_scale, grads->untyped_data(),
                             /*accumulate=*/0, const_cast<void *>(workspace.untyped_data()),
                             static_cast<int64_t>(workspace.size_bytes()), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(tnb_forward_ffi, ForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()   // x [B, L]
                                  .Arg<ffi::AnyBuffer>()   // targets [B, C] or empty
                                  .Arg<ffi::AnyBuffer>()   // packed params
                                  .Ret<ffi::AnyBuffer>()   // loss [B]
                                  .Ret<ffi::AnyBuffer>()   // loss_sum f64[1]
                                  .Ret<ffi::AnyBuffer>()   // out [B, C]
                                  .Ret<ffi::AnyBuffer>()   // out_exp i32[B]
                                  .Ret<ffi::AnyBuffer>()   // workspace u8[...]
                                  .Attr<std::string_view>("problem"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(tnb_backward_ffi, BackwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()   // x
                                  .Arg<ffi::AnyBuffer>()   // targets
                                  .Arg<ffi::AnyBuffer>()   // packed params
                                  .Arg<ffi::AnyBuffer>()   // workspace (residual)
                                  .Ret<ffi::AnyBuffer>()   // packed grads
                                  .Attr<std::string_view>("problem")
                                  .Attr<double>("loss_scale"));

#endif  // TNB_HAVE_XLA_FFI
