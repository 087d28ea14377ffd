// XLA FFI handlers for the FORDE StatefulLayer kernels -- the binding a JAX maintainer would register with
// jax.ffi.register_ffi_target(name, capsule, platform="CUDA") and call through jax.ffi.ffi_call inside a
// jax.custom_vjp.
//
// COMPILE-GATED IN THIS IMAGE: jaxlib (and with it xla/ffi/api/ffi.h) is not installed and cannot be installed here
// (no network), so this file is not part of libforde_b200.so and __graft_entry__.build() skips it.  What IS checked
// here: tests/test_abi.py compiles it with -fsyntax-only against a minimal stub of the FFI types
// (tests/stubs/xla/ffi/api/ffi.h), which verifies the includes, the C-ABI call signatures and that every handler's
// parameter list matches its binding.  It has never run against a real jaxlib.  Build it where jaxlib exists:
//   g++ -O2 -fPIC -shared -DFORDE_WITH_XLA_FFI -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I include -I /usr/local/cuda/include forde-vision-language_b200/csrc/forde_xla_ffi.cc \
//       -L forde-vision-language_b200 -lforde_b200 -L /usr/local/cuda/lib64 -lcudart \
//       -o forde-vision-language_b200/libforde_xla_ffi.so
// It contains no kernel code: every handler forwards plain pointers and sizes to the C ABI of
// include/forde_b200.h, on XLA's compute stream, with buffers XLA allocated.
#ifdef FORDE_WITH_XLA_FFI

#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "forde_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error status(int rc, const char* what) {
  if (rc == FORDE_OK) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + ": " + forde_error_string(rc) + " [" +
                                                   forde_last_cuda_error() + "]");
}

// StatefulLayer forward: (x, w_up, b_up, w_down, b_down, assign) -> (y, zT, hT, act_partials)
static ffi::Error SlFwdImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> x, ffi::Buffer<ffi::BF16> w_up,
                            ffi::Buffer<ffi::F32> b_up, ffi::Buffer<ffi::BF16> w_down, ffi::Buffer<ffi::F32> b_down,
                            ffi::Buffer<ffi::S8> assign, ffi::ResultBuffer<ffi::BF16> y,
                            ffi::ResultBuffer<ffi::BF16> zT, ffi::ResultBuffer<ffi::BF16> hT,
                            ffi::ResultBuffer<ffi::F32> act_partials) {
  const int M = static_cast<int>(x.dimensions()[0]), d = static_cast<int>(x.dimensions()[1]);
  const int H = static_cast<int>(w_up.dimensions()[1]);
  const long long ldz = zT->dimensions()[1];
  return status(forde_sl_fwd(stream, x.untyped_data(), w_up.untyped_data(), b_up.typed_data(), w_down.untyped_data(),
                             b_down.typed_data(), reinterpret_cast<const int8_t*>(assign.untyped_data()),
                             zT->untyped_data(), hT->untyped_data(), y->untyped_data(), act_partials->typed_data(), M,
                             d, H, ldz, /*save_z=*/0),
                "forde_sl_fwd");
}

// StatefulLayer backward: (x, dy, w_up, w_down, zT, hT, assign) -> (dx, dw_up, db_up, dw_down, db_down, grad_partials)
// dw / db results must be zero-initialised by the caller (they are accumulated into); dzT is scratch.
static ffi::Error SlBwdImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> x, ffi::Buffer<ffi::BF16> dy,
                            ffi::Buffer<ffi::BF16> w_up, ffi::Buffer<ffi::BF16> w_down, ffi::Buffer<ffi::BF16> zT,
                            ffi::Buffer<ffi::BF16> hT, ffi::Buffer<ffi::S8> assign, ffi::ResultBuffer<ffi::BF16> dzT,
                            ffi::ResultBuffer<ffi::BF16> dx, ffi::ResultBuffer<ffi::F32> dw_up,
                            ffi::ResultBuffer<ffi::F32> db_up, ffi::ResultBuffer<ffi::F32> dw_down,
                            ffi::ResultBuffer<ffi::F32> db_down, ffi::ResultBuffer<ffi::F32> grad_partials) {
  const int M = static_cast<int>(x.dimensions()[0]), d = static_cast<int>(x.dimensions()[1]);
  const int H = static_cast<int>(w_up.dimensions()[1]);
  const long long ldz = zT.dimensions()[1];
  const cudaError_t zero[4] = {
      cudaMemsetAsync(dw_up->typed_data(), 0, sizeof(float) * d * H, stream),
      cudaMemsetAsync(dw_down->typed_data(), 0, sizeof(float) * d * H, stream),
      cudaMemsetAsync(db_up->typed_data(), 0, sizeof(float) * H, stream),
      cudaMemsetAsync(db_down->typed_data(), 0, sizeof(float) * d, stream)};
  for (cudaError_t e : zero) {
    if (e != cudaSuccess)
      return ffi::Error(ffi::ErrorCode::kInternal, std::string("cudaMemsetAsync: ") + cudaGetErrorString(e));
  }
  return status(forde_sl_bwd(stream, x.untyped_data(), dy.untyped_data(), w_up.untyped_data(), w_down.untyped_data(),
                             zT.untyped_data(), hT.untyped_data(),
                             reinterpret_cast<const int8_t*>(assign.untyped_data()), dzT->untyped_data(),
                             dx->untyped_data(), dw_up->typed_data(), db_up->typed_data(), dw_down->typed_data(),
                             db_down->typed_data(), grad_partials->typed_data(), M, d, H, ldz),
                "forde_sl_bwd");
}

// Sensing finalisation: (act_partials, grad_partials) -> step_stats (H,5); M is a static attribute.
static ffi::Error StatsImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> act_partials,
                            ffi::Buffer<ffi::F32> grad_partials, int32_t M, ffi::ResultBuffer<ffi::F32> step_stats) {
  const int P = static_cast<int>(act_partials.dimensions()[0]);
  const int H = static_cast<int>(act_partials.dimensions()[2]);
  return status(forde_stats_finalize(stream, act_partials.typed_data(), grad_partials.typed_data(), P, H, M,
                                     step_stats->typed_data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr),
                "forde_stats_finalize");
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(forde_sl_fwd_ffi, SlFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(forde_sl_bwd_ffi, SlBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::S8>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(forde_stats_finalize_ffi, StatsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("M")
                                  .Ret<ffi::Buffer<ffi::F32>>());

#endif  // FORDE_WITH_XLA_FFI
