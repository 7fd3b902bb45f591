// orb_plan.h -- per-shape execution plan shared by host code and kernels (POD, passed by value).
#pragma once
#include <stdint.h>

#define ORB_MAX_LEVELS 16
#define ORB_SORT_CAP 8192          // max keypoints kept per (frame, level): select kernel sorts in smem

// one pyramid level of one frame shape
struct OrbLevel {
  int w, h;              // level size: round(W/s), round(H/s), s = (float)pow(f, l)  (ref src/orb_cpu.cpp:284-285)
  int pitch;             // bytes between rows of the u8 level image in the scratch arena (multiple of 16)
  int bpitch;            // elements between rows of the u16 5x5 box-sum image (multiple of 8)
  int tiles_x, tiles_y;  // tile grid of the FAST kernel (128x64 tiles)
  int tile_ofs;          // first flattened tile id of this level inside a frame
  int a_tiles_x, a_tiles_y, a_tile_ofs;   // tile grid of the pyramid kernel (128x32 tiles, levels >= 1)
  int quota;             // keypoints kept on this level (ref src/orb.cpp:62, or nfeatures for raster-first-N)
  int cand_cap;          // candidate slots of this level
  int kept_ofs;          // offset of this level's kept list inside a frame's kept arrays
  int xtab_ofs, ytab_ofs;   // offsets into the resize tap tables
  int edge_ofs, edge_w;     // BRIEF border tables of the level: ey[edge_w] column strips, then rs[h] row sums
  int edge2_ofs;            // border-box value tables of the level (k_edges): bot[2][edge_w], right[2][hh], corner[4]
  float scale;           // (float)pow(f, l): level -> level-0 coordinate scale (ref src/orb.cpp:95)
  unsigned long long lvl_ofs;    // byte offset of the level image inside a frame's pyramid scratch (level 0 unused)
  unsigned long long box_ofs;    // element offset of the box-sum image inside a frame's box scratch
  unsigned long long cand_ofs;   // element offset of the candidate keys inside a frame's candidate scratch
};

struct OrbPlan {
  int nlevels;
  int W, H;                  // level-0 size
  int tiles_per_frame;
  int a_tiles_per_frame;
  int kept_per_frame;        // sum of kept slots over levels
  int edge_frame_elems;      // ints of BRIEF border tables per frame
  int edge2_frame_elems;     // ints of border-box value tables per frame
  int fast_threshold, fast_n, nms_radius, patch_radius;
  int select_policy, blur_levels;
  float harris_k;
  unsigned long long pyr_frame_bytes, box_frame_elems, cand_frame_elems;
  OrbLevel lv[ORB_MAX_LEVELS];
};

// bilinear tap of cv::resize(INTER_LINEAR): source indices and 11-bit weights (sum 2048)
struct OrbTap { uint16_t s0, s1; int16_t a0, a1; };
