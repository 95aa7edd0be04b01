// TEST INFRASTRUCTURE: C entry point around the reference's own HungarianOp (hungarian.cc compiled unmodified against
// the stand-in headers of this directory).  Same contract as oracle/hungarian_ref.c's entry point.
#include "tensorflow/core/framework/op_kernel.h"

::tensorflow::OpKernel *ra_ref_make_kernel();  // defined by REGISTER_KERNEL_BUILDER in hungarian.cc

// W [B,nx,ny] (rank 3) or [nx,ny] (B = 0 selects the rank-2 form).  Returns 0, or -1 when the reference hit one of
// its LOG(FATAL) paths (BFS / max-flow iteration caps, bad rank); msg receives the fatal text.
extern "C" int ref_hungarian_f32(const float *W, int B, int nx, int ny, float *M, float *cx, float *cy, char *msg,
                                 int msg_len) {
  using namespace tensorflow;
  TensorShape s;
  const int nb = B > 0 ? B : 1;
  if (B > 0) s.AddDim(B);
  s.AddDim(nx);
  s.AddDim(ny);
  OpKernelContext ctx;
  ctx.inputs.push_back(Tensor(s));
  std::memcpy(ctx.inputs[0].raw(), W, sizeof(float) * (size_t)nb * nx * ny);
  OpKernel *k = ra_ref_make_kernel();
  int rc = 0;
  std::streambuf *cout_buf = std::cout.rdbuf(nullptr);  // hungarian.cc prints its S / T sets to stdout (PrintSet)
  try {
    k->Compute(&ctx);
    std::memcpy(M, ctx.outputs[0].raw(), sizeof(float) * (size_t)nb * nx * ny);
    std::memcpy(cx, ctx.outputs[1].raw(), sizeof(float) * (size_t)nb * nx);
    std::memcpy(cy, ctx.outputs[2].raw(), sizeof(float) * (size_t)nb * ny);
  } catch (const FatalError &e) {
    rc = -1;
    if (msg && msg_len > 0) {
      std::strncpy(msg, e.what(), (size_t)msg_len - 1);
      msg[msg_len - 1] = 0;
    }
  }
  std::cout.rdbuf(cout_buf);
  std::cout.clear();
  delete k;
  return rc;
}
