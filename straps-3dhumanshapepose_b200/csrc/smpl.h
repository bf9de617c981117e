// smpl.h -- constants and the device-side model description shared by smpl.cu (forward) and smpl_bwd.cu (backward).
#pragma once
#include "common.cuh"
#include "../../include/straps_b200.h"
#include <vector>

namespace straps {

constexpr int V = STRAPS_NUM_VERTS;
constexpr int TV = 128;                 // vertices per CTA
constexpr int NTILES = (V + TV - 1) / TV;   // 54
constexpr int VPAD = NTILES * TV;       // 6912
constexpr int VP3 = VPAD * 3;
constexpr int NJ = STRAPS_NUM_JOINTS;
constexpr int NPF = (NJ - 1) * 9;       // 207
constexpr int NPF_PAD = 208;
constexpr int KC = 8;                   // posedirs rows per pipeline stage
constexpr int NSTAGE = 3;
constexpr int NCHUNK = NPF_PAD / KC;    // 26
constexpr int ROWF = TV * 3;            // 384 floats per slab row

struct SmplDev {
  const float* vt;
  const float* sdir;
  const float* pdir;
  const float* jt;
  const float* js;
  const int* widx;
  const float* wval;
  const float* wdense;
  const int* pick_idx;
  const int* csr_ptr;
  const int* csr_idx;
  const float* csr_val;
  int parents[NJ];
  int lvl_joint[NJ];      // joints sorted by tree depth
  int lvl_start[NJ + 1];  // level l owns lvl_joint[lvl_start[l] .. lvl_start[l+1])
  int nlevels;
};

}  // namespace straps

struct straps_smpl {
  straps::SmplDev d;
  int sparse4;
  std::vector<void*> allocs;
  // tensor-core LBS (smpl_tc.cu): [posedirs ; shapedirs ; v_template] packed once as swizzled fp16 hi / lo images + row unscales,
  // and a per-call scratch (Bm images + skinning transforms) that grows with the largest batch seen
  const unsigned char* tc_apk;
  const float* tc_ainv;
  const unsigned char* tc_wk;
  unsigned char* tc_at_base;
  void* tc_scratch;
  size_t tc_scratch_bytes;
};

namespace straps {
int smpl_tc_pack(const float* v_template, const float* shapedirs, const float* posedirs, const float* lbs_weights,
                 std::vector<unsigned char>& apk, std::vector<float>& ainv, std::vector<unsigned char>& wk);
int smpl_tc_forward(straps_smpl* m, const float* global_orient, int64_t go_stride, const float* body_pose, int64_t bp_stride,
                    const float* betas, int64_t betas_stride, const float* transl, int batch, int pose2rot, float* vertices,
                    float* joints, float* save_vposed, float* save_A, cudaStream_t st);
}  // namespace straps
