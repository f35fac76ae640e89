// stacb_xla_ffi.cc -- XLA FFI handlers (typed FFI API, jaxlib >= 0.4.31) around the C ABI of include/stacb.h.
//
// NOT BUILT IN THE AUTHORING IMAGE: jax / jaxlib (and therefore xla/ffi/api/ffi.h) are not installed there.  The translation unit is
// type-checked (g++ -fsyntax-only) against a mock of the FFI API's shape (tests/xla_ffi_mock, tests/test_abi_cpu.py): every handler
// matches its binding and every C ABI call matches include/stacb.h; it has never run inside XLA.  build.sh builds it into
// libstacb_xla_ffi.so only when
// `python -c "import jax.ffi; print(jax.ffi.include_dir())"` succeeds.  It contains no arithmetic: every handler unpacks
// XLA buffers into the plain pointers of the C ABI and forwards the stream XLA hands it.  Python side: stac_mjx_b200/jax_ffi.py.
//
// Replaces, for a JAX host, reference stac_mjx/stac_core.py:66-99 (_q_opt jit) and the vmapped drivers of
// stac_mjx/stac.py:405-440.
#include <cstdint>

#include <cuda_runtime.h>

#include "xla/ffi/api/ffi.h"

#include "stacb.h"

namespace ffi = xla::ffi;

static ffi::Error to_error(int rc) {
  if (rc == STACB_OK) return ffi::Error::Success();
  return ffi::Error(rc == STACB_E_INVALID ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal, stacb_last_error());
}

// Fused root_optimization + pose_optimization over clips.  Attribute `tree` is the stacb_tree* handle as int64.
static ffi::Error PoseClipsImpl(cudaStream_t stream, int64_t tree, int32_t do_root, int32_t root_kp_idx, int32_t root_dims, float tol,
                                int32_t maxiter, int32_t maxls,
                                ffi::Buffer<ffi::F32> kp,          // [C, F, 3K]
                                ffi::Buffer<ffi::F32> qpos_init,   // [C, nq]
                                ffi::Buffer<ffi::F32> site_pos,    // [K, 3]
                                ffi::Buffer<ffi::F32> lb, ffi::Buffer<ffi::F32> ub,
                                ffi::Buffer<ffi::U8> part_masks,   // [P, nq]
                                ffi::Buffer<ffi::U8> trunk_kps,    // [K]
                                ffi::ResultBuffer<ffi::F32> qpos,       // [C, F, nq]
                                ffi::ResultBuffer<ffi::F32> xpos,       // [C, F, nbody, 3]
                                ffi::ResultBuffer<ffi::F32> xquat,      // [C, F, nbody, 4]
                                ffi::ResultBuffer<ffi::F32> sites,      // [C, F, K, 3]
                                ffi::ResultBuffer<ffi::F32> err,        // [C, F]
                                ffi::ResultBuffer<ffi::F32> qpos_last,  // [C, nq]  (qpos_io of the C ABI)
                                ffi::ResultBuffer<ffi::S32> iters,      // [C, F, 1 + P]
                                ffi::ResultBuffer<ffi::S32> ls_evals,   // [C, F, 1 + P]
                                ffi::ResultBuffer<ffi::S32> root_stats, // [C, 4]
                                ffi::ResultBuffer<ffi::S32> status) {   // [C]
  const auto dims = kp.dimensions();
  if (dims.size() != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "kp must be [C, F, 3K]");
  const int C = static_cast<int>(dims[0]), F = static_cast<int>(dims[1]);
  const int P = part_masks.dimensions().size() == 2 ? static_cast<int>(part_masks.dimensions()[0]) : 0;
  // the C ABI updates qpos_io in place: seed the output buffer with the warm start
  cudaError_t ce = cudaMemcpyAsync(qpos_last->typed_data(), qpos_init.typed_data(), qpos_init.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  if (ce != cudaSuccess) return ffi::Error(ffi::ErrorCode::kInternal, cudaGetErrorString(ce));
  return to_error(stacb_pose_clips(reinterpret_cast<const stacb_tree *>(tree), kp.typed_data(), qpos_last->typed_data(), site_pos.typed_data(),
                                   lb.typed_data(), ub.typed_data(), P ? part_masks.typed_data() : nullptr, P, do_root, root_kp_idx,
                                   trunk_kps.typed_data(), root_dims, tol, maxiter, maxls, qpos->typed_data(), xpos->typed_data(),
                                   xquat->typed_data(), sites->typed_data(), err->typed_data(), iters->typed_data(), ls_evals->typed_data(),
                                   root_stats->typed_data(), status->typed_data(), C, F, stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(StacbPoseClips, PoseClipsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("tree")
                                  .Attr<int32_t>("do_root")
                                  .Attr<int32_t>("root_kp_idx")
                                  .Attr<int32_t>("root_dims")
                                  .Attr<float>("tol")
                                  .Attr<int32_t>("maxiter")
                                  .Attr<int32_t>("maxls")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

// stac_core._q_opt for a batch of independent solves.
static ffi::Error QOptImpl(cudaStream_t stream, int64_t tree, float tol, int32_t maxiter, int32_t maxls, ffi::Buffer<ffi::F32> q0,
                           ffi::Buffer<ffi::F32> kp, ffi::Buffer<ffi::U8> q_mask, ffi::Buffer<ffi::U8> kp_mask, ffi::Buffer<ffi::F32> site_pos,
                           ffi::Buffer<ffi::F32> lb, ffi::Buffer<ffi::F32> ub, ffi::ResultBuffer<ffi::F32> params,
                           ffi::ResultBuffer<ffi::F32> error, ffi::ResultBuffer<ffi::S32> iters, ffi::ResultBuffer<ffi::S32> ls_evals) {
  const int B = static_cast<int>(q0.dimensions()[0]);
  return to_error(stacb_q_opt(reinterpret_cast<const stacb_tree *>(tree), q0.typed_data(), kp.typed_data(), q_mask.typed_data(),
                              kp_mask.typed_data(), site_pos.typed_data(), lb.typed_data(), ub.typed_data(), tol, maxiter, maxls,
                              params->typed_data(), error->typed_data(), iters->typed_data(), ls_evals->typed_data(), B, stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(StacbQOpt, QOptImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("tree")
                                  .Attr<float>("tol")
                                  .Attr<int32_t>("maxiter")
                                  .Attr<int32_t>("maxls")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

// utils.kinematics + get_site_xpos.
static ffi::Error FkImpl(cudaStream_t stream, int64_t tree, ffi::Buffer<ffi::F32> qpos, ffi::Buffer<ffi::F32> site_pos,
                         ffi::ResultBuffer<ffi::F32> qpos_out, ffi::ResultBuffer<ffi::F32> xpos, ffi::ResultBuffer<ffi::F32> xquat,
                         ffi::ResultBuffer<ffi::F32> site_xpos) {
  const int B = static_cast<int>(qpos.dimensions()[0]);
  return to_error(stacb_fk(reinterpret_cast<const stacb_tree *>(tree), qpos.typed_data(), site_pos.typed_data(), qpos_out->typed_data(),
                           xpos->typed_data(), xquat->typed_data(), site_xpos->typed_data(), B, stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(StacbFk, FkImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("tree")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// _m_opt sufficient statistics; `scratch` is an extra result so XLA owns the temporary.
// out [3K+2] = { s, z2, T }
static ffi::Error MStatsImpl(cudaStream_t stream, int64_t tree, ffi::Buffer<ffi::F32> kp, ffi::Buffer<ffi::F32> q,
                             ffi::ResultBuffer<ffi::F32> out, ffi::ResultBuffer<ffi::F32> scratch) {
  const int T = static_cast<int>(kp.dimensions()[0]);
  return to_error(stacb_m_stats(reinterpret_cast<const stacb_tree *>(tree), kp.typed_data(), q.typed_data(), scratch->typed_data(),
                                out->typed_data(), T, stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(StacbMStats, MStatsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("tree")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());
