// Binary layer plan shared between floodsr_b200/graph.py (writer) and the CUDA engine (reader).
//
// The plan is what `floodsr_b200.graph.lower_onnx` makes of model_infer.onnx: a topologically ordered
// list of fused layer ops over per-tile NHWC activation tensors.  It replaces the ONNX Runtime session
// object of the reference (floodsr/engine/ort.py:54).  All integers are little-endian int32.
//
//   header  : FSR_PLAN_MAGIC, version, n_tensors, n_ops, lr_tile, hr_tile, scale, out_tensor
//   tensors : n_tensors x { h, w, c }              tensor 0 = depth_lr (lr,lr,1), tensor 1 = dem_hr (hr,hr,1)
//   ops     : n_ops x fsr_op                       (fixed 16 x int32 records)
//
// Weights live in one float32 blob; conv kernels are stored [kh][kw][cin][cout] (HWIO), transposed-conv
// kernels [kh][kw][cin][cout], biases [cout].  Offsets are in floats, -1 = absent.
#pragma once
#include <stdint.h>

#define FSR_PLAN_MAGIC 0x50525346 /* 'FSRP' */
#define FSR_PLAN_VERSION 2

enum fsr_op_kind {
  FSR_OP_CONV = 1,      // same-padded k x k conv, stride 1, over concat(src0, src1); + bias + residual, act
  FSR_OP_POOL = 2,      // k x k pooling with stride k (mode: 0 max, 1 average, 2 pick element (aux, aux) of every block: the
                        // sampling of a strided convolution, which is lowered as stride-1 conv + pick)
  FSR_OP_UPSAMPLE = 3,  // upsampling by integer factor k (mode: 0 nearest, 1 / 2 / 3 bilinear with ONNX Resize's half_pixel /
                        // align_corners / asymmetric coordinate transformation)
  FSR_OP_CONVT = 4,     // transposed conv with kernel == stride == k (non-overlapping patches); + bias, act
  FSR_OP_ELTWISE = 5,   // dst = act(src0 + src1)   (src1 may be -1); the only op that takes FSR_ACT_CLIP / FSR_ACT_SIGMOID
  FSR_OP_HEAD = 6,      // fused head: y = act(conv kxk(concat(src0, src1)) + b); out = conv1x1(y) (+ b2), linear
};

enum fsr_act { FSR_ACT_NONE = 0, FSR_ACT_RELU = 1, FSR_ACT_LEAKY = 2, FSR_ACT_CLIP = 3 /* [alpha, beta] */, FSR_ACT_SIGMOID = 4 };
enum fsr_pool_mode { FSR_POOL_MAX = 0, FSR_POOL_AVG = 1, FSR_POOL_PICK = 2 };
enum fsr_up_mode { FSR_UP_NEAREST = 0, FSR_UP_LINEAR_HALF_PIXEL = 1, FSR_UP_LINEAR_ALIGN_CORNERS = 2, FSR_UP_LINEAR_ASYMMETRIC = 3 };

struct fsr_op {
  int32_t kind;
  int32_t src0, src1;  // tensor ids (-1 = none)
  int32_t res;         // residual tensor added before the activation (-1 = none)
  int32_t dst;
  int32_t k;           // kernel size / pool size / upsample factor
  int32_t mode;        // pool mode
  int32_t cout;        // conv/convT output channels (HEAD: mid channels)
  int32_t act;
  float alpha;         // leaky-relu slope / clip lower bound
  int32_t w_off, b_off;    // main weights / bias
  int32_t w2_off, b2_off;  // HEAD: 1x1 weights [cmid] / bias
  float beta;          // clip upper bound
  int32_t aux;         // FSR_POOL_PICK: offset of the picked element inside its block
};

struct fsr_plan_header {
  int32_t magic, version, n_tensors, n_ops, lr_tile, hr_tile, scale, out_tensor;
};

struct fsr_tensor_desc {
  int32_t h, w, c;
};
