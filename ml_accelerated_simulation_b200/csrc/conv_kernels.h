// Launchers of the correction-CNN kernels (definitions in conv_kernels.cu).
#pragma once
#include "common.cuh"

namespace mlsim {

struct ConvFirstArgs {
  const float* u; const float* v;     // fp32 state [batch][nx][ny]
  __nv_bfloat16* out;                 // activation buffer act[b][x][8][ny+2][8]
  const uint8_t* wimg;                // packed weights (conv_first_wimg_bytes)
  float scale;                        // cnn_input_scale
  int nx, ny, batch;
  long long img_stride;               // elements between consecutive images of u / v (nx*ny unless halo-extended)
  int reverse;                        // traverse tiles last-to-first (see ConvMidArgs::reverse)
};

struct ConvMidArgs {
  const __nv_bfloat16* in;            // activation buffer
  __nv_bfloat16* out;                 // mid layers: activation buffer
  float* u; float* v;                 // last layer, fused correction add (when y_nchw == nullptr)
  float* y_nchw;                      // last layer, test hook output [batch][cout][nx][ny]
  const uint8_t* wimg;                // packed weights + bias
  int nx, ny, batch;
  long long img_stride;               // elements between consecutive images of u / v (fused correction add)
  int rs;                             // output rows per work item (strip length)
  int cout;                           // real output channels of this layer
  int relu;
  int reverse;                        // traverse work items last-to-first: consecutive layers alternate direction so
                                      // each layer starts on the activations its predecessor wrote last (still in L2)
  unsigned long long* dbg;            // optional [grid][8] cycle counters (nullptr = off): see mlsim_conv_debug
  // shortcut added before the activation (ResNet blocks, reference src/models/ResNet.py:42-50: act(x + y)):
  //   0 none
  //   1 mid layers: identity, `res` = block input in the activation layout
  //   2 mid layers: 1x1 convolution of the 2-channel network input (u, v) * sscale; weights [64][2] fp32 after the bias
  //   3 last layer: 1x1 convolution (64 -> cout) of the block input `res`; weights [16][64] fp32 after the bias
  int res_mode;
  const __nv_bfloat16* res;
  float sscale;
  // fused first layer (launch_conv_mid_fused): the kernel's input rows are not read from `in` but COMPUTED from the fp32
  // state (fu, fv) * fscale by the first layer (2 -> 64, weights wimg0 = the first layer's operand image)
  const float* fu; const float* fv;
  const uint8_t* wimg0;
  float fscale;
};

int launch_conv_first(const ConvFirstArgs& a, int sm_count, cudaStream_t st);
int launch_conv_mid(const ConvMidArgs& a, int sm_count, cudaStream_t st);
int launch_conv_mid_fused(const ConvMidArgs& a, int sm_count, cudaStream_t st);     // first layer + first hidden layer
int launch_conv_last(const ConvMidArgs& a, int sm_count, cudaStream_t st);
int conv_pick_strip(int nx, int ny, int batch, int sm_count);
int conv_first_wimg_bytes();
int conv_mid_wimg_bytes();        // [operand image][bias 64 f32][1x1 shortcut weights 64x2 f32]
int conv_last_wimg_bytes();       // [operand image][bias 16 f32][1x1 shortcut weights 16x64 f32]
int conv_mid_extra_off();         // byte offset of the shortcut weights inside the image
int conv_last_extra_off();

}  // namespace mlsim
