// Pixel-tile geometry shared by the conv kernel (conv_igemm.cu) and the GroupNorm finalize kernel
// (norm.cu): a GEMM M tile is TN x TH x TW output pixels with TN*TH*TW = 128, all powers of two.
#pragma once
namespace srgd {
struct TileGeom {
  int tw_log2, th_log2, tn_log2, tiles_x, tiles_y, tiles_b, m_tiles;
};
static inline int ceil_log2_i(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}
static inline TileGeom tile_geom(int B, int Ho, int Wo) {
  TileGeom g;
  g.tw_log2 = ceil_log2_i(Wo) < 7 ? ceil_log2_i(Wo) : 7;
  int rem = 7 - g.tw_log2;
  g.th_log2 = ceil_log2_i(Ho) < rem ? ceil_log2_i(Ho) : rem;
  g.tn_log2 = 7 - g.tw_log2 - g.th_log2;
  g.tiles_x = (Wo + (1 << g.tw_log2) - 1) >> g.tw_log2;
  g.tiles_y = (Ho + (1 << g.th_log2) - 1) >> g.th_log2;
  g.tiles_b = (B + (1 << g.tn_log2) - 1) >> g.tn_log2;
  g.m_tiles = g.tiles_x * g.tiles_y * g.tiles_b;
  return g;
}
}  // namespace srgd
