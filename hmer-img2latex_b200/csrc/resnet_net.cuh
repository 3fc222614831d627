// torchvision ResNet trunk topology (model/encoder.py:184-199 builds it from torchvision.models.resnetNN minus fc)
// shared by the fp32 path (encoder.cu) and the bf16 tcgen05 path (resnet_bf16.cu).
#pragma once
#include "common.cuh"
#include <algorithm>
#include <vector>

namespace i2l {

struct RConv { int ci, co, k, stride, pad; size_t w_off, b_off; };
struct RBlock { int c1, c2, c3, ds; };
struct RNet { std::vector<RConv> convs; std::vector<RBlock> blocks; int feat; bool ok; };

inline RNet build_resnet(int depth) {
  RNet n; n.ok = true; n.feat = 0;
  bool bottleneck; int layers[4];
  switch (depth) {
    case 18: bottleneck = false; layers[0] = 2; layers[1] = 2; layers[2] = 2; layers[3] = 2; break;
    case 34: bottleneck = false; layers[0] = 3; layers[1] = 4; layers[2] = 6; layers[3] = 3; break;
    case 50: bottleneck = true; layers[0] = 3; layers[1] = 4; layers[2] = 6; layers[3] = 3; break;
    case 101: bottleneck = true; layers[0] = 3; layers[1] = 4; layers[2] = 23; layers[3] = 3; break;
    case 152: bottleneck = true; layers[0] = 3; layers[1] = 8; layers[2] = 36; layers[3] = 3; break;
    default: n.ok = false; return n;
  }
  auto add = [&](int ci, int co, int k, int s, int p) { n.convs.push_back(RConv{ci, co, k, s, p, 0, 0}); return (int)n.convs.size() - 1; };
  add(3, 64, 7, 2, 3);
  int inpl = 64, exp = bottleneck ? 4 : 1;
  for (int li = 0; li < 4; ++li) {
    int planes = 64 << li;
    for (int b = 0; b < layers[li]; ++b) {
      int stride = (li > 0 && b == 0) ? 2 : 1;
      RBlock blk{-1, -1, -1, -1};
      if (!bottleneck) {
        blk.c1 = add(inpl, planes, 3, stride, 1);
        blk.c2 = add(planes, planes, 3, 1, 1);
      } else {
        blk.c1 = add(inpl, planes, 1, 1, 0);
        blk.c2 = add(planes, planes, 3, stride, 1);
        blk.c3 = add(planes, planes * 4, 1, 1, 0);
      }
      if (stride != 1 || inpl != planes * exp) blk.ds = add(inpl, planes * exp, 1, stride, 0);
      inpl = planes * exp;
      n.blocks.push_back(blk);
    }
  }
  n.feat = inpl;
  size_t o = 0;
  auto take = [&](size_t c) { size_t r = o; o += (c + 63) / 64 * 64; return r; };
  for (auto& c : n.convs) { c.w_off = take((size_t)c.co * c.ci * c.k * c.k); c.b_off = take(c.co); }
  return n;
}

struct RLayout { size_t fc_w, fc_b, end_f32; };
inline RLayout resnet_layout(const RNet& n, int E) {
  size_t o = n.convs.back().b_off + (n.convs.back().co + 63) / 64 * 64;
  RLayout L{};
  L.fc_w = o; o += ((size_t)E * n.feat + 63) / 64 * 64;
  L.fc_b = o; o += (E + 63) / 64 * 64;
  L.end_f32 = o;
  return L;
}

inline size_t resnet_max_act(const RNet& n, int H, int W) {
  // per-image float count of the largest activation
  auto o = [](int x, int k, int s, int p) { return (x + 2 * p - k) / s + 1; };
  int h = o(H, 7, 2, 3), w = o(W, 7, 2, 3);
  size_t mx = (size_t)64 * h * w;
  h = o(h, 3, 2, 1); w = o(w, 3, 2, 1);
  for (auto& b : n.blocks) {
    const RConv& c1 = n.convs[b.c1];
    const RConv& c2 = n.convs[b.c2];
    int h1 = o(h, c1.k, c1.stride, c1.pad), w1 = o(w, c1.k, c1.stride, c1.pad);
    mx = std::max(mx, (size_t)c1.co * h1 * w1);
    int h2 = o(h1, c2.k, c2.stride, c2.pad), w2 = o(w1, c2.k, c2.stride, c2.pad);
    mx = std::max(mx, (size_t)c2.co * h2 * w2);
    if (b.c3 >= 0) mx = std::max(mx, (size_t)n.convs[b.c3].co * h2 * w2);
    h = h2; w = w2;
  }
  return mx;
}


}  // namespace i2l
