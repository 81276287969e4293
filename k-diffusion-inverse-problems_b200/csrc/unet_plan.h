// Block plan of the ADM UNet (the walk of UNetModel.__init__, guided_diffusion/unet.py:482-618), shared by the bf16 tcgen05
// engine (unet.cu) and the fp32 reference-precision engine (unet_fp32.cu).
#pragma once
#include <string>
#include <vector>

#include "../../include/kdip.h"

namespace kdip {

struct BlockDesc {
  std::string prefix;
  int kind;      // 0 conv_in, 1 res, 2 attn
  int cin, cout;
  int updown;    // 0 none, 1 down, 2 up
  int stage;     // 0 in, 1 mid, 2 out
  int block;     // index of the enclosing TimestepEmbedSequential
  int skip_ch;   // channels popped from the skip stack (out-stage res blocks that start a block)
  bool first_of_block, last_of_block;
};

void build_block_plan(const kdip_unet_arch& a, std::vector<BlockDesc>& plan);

}  // namespace kdip
