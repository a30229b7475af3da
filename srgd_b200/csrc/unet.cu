// Host-side launcher for the whole denoiser: ConditionalSRUnet.forward (model.py:678-725) expressed
// as a fixed sequence of srgd_b200 kernels over a packed parameter list and one caller-provided
// workspace.  No allocation, no synchronisation: buffer addresses come from a deterministic
// first-fit arena over the workspace, so the same (B,H,W) always yields the same launch sequence
// with the same pointers (CUDA-graph friendly).
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "tiles.h"

namespace srgd {

// ---------------------------------------------------------------------------------------------
// workspace arena
// ---------------------------------------------------------------------------------------------
class Arena {
 public:
  Arena(void* base, size_t cap) : base_(reinterpret_cast<uint8_t*>(base)), cap_(cap) {
    blks_.push_back({0, cap, true});
  }
  // returns nullptr when exhausted (and records it)
  void* alloc(size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    if (bytes == 0) bytes = 256;
    for (size_t i = 0; i < blks_.size(); ++i) {
      if (blks_[i].free && blks_[i].size >= bytes) {
        if (blks_[i].size > bytes) {
          Blk rest{blks_[i].off + bytes, blks_[i].size - bytes, true};
          blks_[i].size = bytes;
          blks_.insert(blks_.begin() + i + 1, rest);
        }
        blks_[i].free = false;
        if (blks_[i].off + bytes > high_) high_ = blks_[i].off + bytes;
        return base_ + blks_[i].off;
      }
    }
    failed_ = true;
    return nullptr;
  }
  void release(void* p) {
    if (p == nullptr) return;
    const size_t off = reinterpret_cast<uint8_t*>(p) - base_;
    for (size_t i = 0; i < blks_.size(); ++i) {
      if (blks_[i].off == off && !blks_[i].free) {
        blks_[i].free = true;
        if (i + 1 < blks_.size() && blks_[i + 1].free) {
          blks_[i].size += blks_[i + 1].size;
          blks_.erase(blks_.begin() + i + 1);
        }
        if (i > 0 && blks_[i - 1].free) {
          blks_[i - 1].size += blks_[i].size;
          blks_.erase(blks_.begin() + i);
        }
        return;
      }
    }
  }
  size_t high_water() const { return high_; }
  bool failed() const { return failed_; }

 private:
  struct Blk {
    size_t off, size;
    bool free;
  };
  uint8_t* base_;
  size_t cap_;
  size_t high_ = 0;
  bool failed_ = false;
  std::vector<Blk> blks_;
};

// ---------------------------------------------------------------------------------------------
// parameter bookkeeping
// ---------------------------------------------------------------------------------------------
struct ResP {
  std::string name;
  int cin0, cin1, cout, ss_off;
  const void *c1_w, *c2_w, *res_w;
  const float *c1_b, *n1_g, *n1_b, *c2_b, *n2_g, *n2_b, *res_b;
};
struct AttnP {
  std::string name;
  int C;
  bool full;
  const void *qkv_w, *out_w;
  const float *out_b, *out_g;
};
struct ConvP {
  std::string name;
  int cin, cout;
  const void* w;
  const float* b;
};
struct Stage {
  ResP r0, r1;
  AttnP attn;
  ConvP resample;      // Downsample / PixelShuffleUpsample 1x1, or the last stage's conv3x3
  bool last;
};

struct Tap {
  std::string name;
  void* dst;
  size_t bytes;
};

}  // namespace srgd

struct srgd_unet {
  srgd_unet_config cfg;
  std::vector<std::string> names;
  std::vector<const void*> ptrs;
  int time_dim, hidden, ss_total;
  const void* init_w;
  const float *init_b, *time_freq, *time_w1, *time_b1, *time_w2, *time_b2, *class_table, *ss_w, *ss_b, *final_w,
      *final_b;
  std::vector<srgd::Stage> downs, ups;
  srgd::ResP mid1, mid2, final_res;
  srgd::AttnP mid_attn;
  std::vector<srgd::Tap> taps;
  long last_launches = 0;
};

namespace srgd {

// Walks the architecture in forward order; either collects names (u.ptrs empty) or binds pointers.
struct Binder {
  srgd_unet& u;
  bool bind;
  size_t cursor = 0;
  const void* next(const std::string& name) {
    if (!bind) {
      u.names.push_back(name);
      return nullptr;
    }
    return u.ptrs[cursor++];
  }
  const float* nextf(const std::string& name) { return reinterpret_cast<const float*>(next(name)); }
};

static void walk_res(Binder& b, ResP& r, const std::string& name, int cin0, int cin1, int cout, int& ss_off) {
  r.name = name;
  r.cin0 = cin0; r.cin1 = cin1; r.cout = cout;
  r.ss_off = ss_off;
  ss_off += 2 * cout;
  r.c1_w = b.next(name + ".c1.w"); r.c1_b = b.nextf(name + ".c1.b");
  r.n1_g = b.nextf(name + ".n1.g"); r.n1_b = b.nextf(name + ".n1.b");
  r.c2_w = b.next(name + ".c2.w"); r.c2_b = b.nextf(name + ".c2.b");
  r.n2_g = b.nextf(name + ".n2.g"); r.n2_b = b.nextf(name + ".n2.b");
  if (cin0 + cin1 != cout) {
    r.res_w = b.next(name + ".res.w"); r.res_b = b.nextf(name + ".res.b");
  } else {
    r.res_w = nullptr; r.res_b = nullptr;
  }
}
static void walk_attn(Binder& b, AttnP& a, const std::string& name, int C, bool full) {
  a.name = name; a.C = C; a.full = full;
  a.qkv_w = b.next(name + ".qkv.w");
  a.out_w = b.next(name + ".out.w");
  a.out_b = b.nextf(name + ".out.b");
  a.out_g = full ? nullptr : b.nextf(name + ".out.g");
}
static void walk_conv(Binder& b, ConvP& c, const std::string& name, int cin, int cout) {
  c.name = name; c.cin = cin; c.cout = cout;
  c.w = b.next(name + ".w");
  c.b = b.nextf(name + ".b");
}

static void walk(srgd_unet& u, bool bind) {
  Binder b{u, bind};
  const srgd_unet_config& c = u.cfg;
  const int n = c.n_stages;
  std::vector<int> dims(n + 1);
  dims[0] = c.dim;
  for (int i = 0; i < n; ++i) dims[i + 1] = c.dim * c.dim_mults[i];
  u.time_dim = c.dim * 4;
  u.hidden = c.heads * c.dim_head;
  int ss_off = 0;
  u.init_w = b.next("init.w"); u.init_b = b.nextf("init.b");
  u.time_freq = b.nextf("time.freq");
  u.time_w1 = b.nextf("time.w1"); u.time_b1 = b.nextf("time.b1");
  u.time_w2 = b.nextf("time.w2"); u.time_b2 = b.nextf("time.b2");
  u.class_table = c.num_classes > 0 ? b.nextf("class.table") : nullptr;
  u.ss_w = b.nextf("ss.w"); u.ss_b = b.nextf("ss.b");
  u.downs.resize(n);
  u.ups.resize(n);
  for (int i = 0; i < n; ++i) {                                       // model.py:638-648
    Stage& s = u.downs[i];
    const int di = dims[i], dout = dims[i + 1];
    const std::string p = "downs." + std::to_string(i);
    walk_res(b, s.r0, p + ".0", di, 0, di, ss_off);
    walk_res(b, s.r1, p + ".1", di, 0, di, ss_off);
    walk_attn(b, s.attn, p + ".2", di, c.full_attn[i] != 0);
    s.last = (i == n - 1);
    walk_conv(b, s.resample, p + ".3", s.last ? di : 4 * di, dout);
  }
  const int mid = dims[n];
  walk_res(b, u.mid1, "mid_block1", mid, 0, mid, ss_off);             // model.py:651-653
  walk_attn(b, u.mid_attn, "mid_attn", mid, true);
  walk_res(b, u.mid2, "mid_block2", mid, 0, mid, ss_off);
  for (int i = 0; i < n; ++i) {                                       // model.py:659-669
    Stage& s = u.ups[i];
    const int di = dims[n - 1 - i], dout = dims[n - i];               // (dim_in, dim_out) reversed
    const std::string p = "ups." + std::to_string(i);
    walk_res(b, s.r0, p + ".0", dout, di, dout, ss_off);
    walk_res(b, s.r1, p + ".1", dout, di, dout, ss_off);
    walk_attn(b, s.attn, p + ".2", dout, c.full_attn[n - 1 - i] != 0);
    s.last = (i == n - 1);
    walk_conv(b, s.resample, p + ".3", dout, s.last ? di : 4 * di);
  }
  walk_res(b, u.final_res, "final_res_block", c.dim, c.dim, c.dim, ss_off);   // model.py:674
  u.final_w = b.nextf("final.w"); u.final_b = b.nextf("final.b");
  u.ss_total = ss_off;
}

static int validate_cfg(const srgd_unet_config* c) {
  SRGD_REQUIRE(c != nullptr, "unet: null config");
  SRGD_REQUIRE(c->dim >= 64 && c->dim % 64 == 0, "unet: dim=%d must be a multiple of 64", c->dim);
  SRGD_REQUIRE(c->n_stages >= 1 && c->n_stages <= 6, "unet: n_stages=%d out of range", c->n_stages);
  SRGD_REQUIRE(c->heads == 4 && c->dim_head == 32, "unet: only heads=4, dim_head=32 is built");
  SRGD_REQUIRE(c->groups == 8, "unet: only resnet_block_groups=8 is built");
  SRGD_REQUIRE(c->channels == 3, "unet: only channels=3 is built");
  SRGD_REQUIRE(c->fixed_sinusoidal == 0 || c->fixed_sinusoidal == 1, "unet: bad fixed_sinusoidal flag");
  SRGD_REQUIRE(c->fixed_sinusoidal || (c->sinu_dim > 0 && c->sinu_dim % 2 == 0), "unet: bad learned_sinusoidal_dim");
  for (int i = 0; i < c->n_stages; ++i) SRGD_REQUIRE(c->dim_mults[i] >= 1, "unet: bad dim_mults");
  return SRGD_OK;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct Fwd {
  srgd_unet& u;
  Arena& ar;
  bool dry;                 // plan only: run the arena, launch nothing
  int conv_impl;            // bit0: CUDA-core direct conv; bit1: stand-alone GroupNorm statistics
  srgd_stream_t st;
  int B;
  int rc = SRGD_OK;
  const float* ss = nullptr;       // [B][ss_total] scale/shift for all ResnetBlocks
  void* sk_ws = nullptr;           // split-K workspace shared by every conv launch of the forward (conv_igemm.cu)
  int64_t sk_bytes = 0;

  bool ok() const { return rc == SRGD_OK; }
  void run(int r) {
    if (rc == SRGD_OK && r != SRGD_OK) rc = r;
  }
  void* alloc(size_t bytes) {
    void* p = ar.alloc(bytes);
    if (p == nullptr && rc == SRGD_OK) {
      set_error("unet_forward: workspace too small");
      rc = SRGD_E_WORKSPACE;
    }
    return p;
  }
  void tap(const std::string& name, const void* src, size_t bytes) {
    if (dry || !ok()) return;
    for (const Tap& t : u.taps)
      if (t.name == name) {
        const size_t nb = bytes < t.bytes ? bytes : t.bytes;
        cudaError_t e = cudaMemcpyAsync(t.dst, src, nb, cudaMemcpyDeviceToDevice, as_stream(st));
        if (e != cudaSuccess) run(fail_cuda(e, "tap copy"));
      }
  }

  // generic conv over up to two concatenated NHWC sources with a kxk window
  // `rows` < 0: all B rows; otherwise a launch over `rows` samples (a group of a broadcast-batch consumer)
  void conv(const void* a0, int c0, const void* a1, int c1, int H, int W, int ksize, const void* w, const float* bias,
            int cout, void* out, float* partials, const float* row_scale, const void* residual, int act, int out_mode,
            int rows = -1) {
    if (dry || !ok()) return;
    srgd_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.B = rows < 0 ? B : rows; d.Ho = H; d.Wo = W; d.Cout = cout;
    d.n_src = a1 ? 2 : 1;
    d.srcs[0] = {a0, (int64_t)H * W * c0, (int64_t)W * c0, (int64_t)c0, H, W, c0};
    if (a1) d.srcs[1] = {a1, (int64_t)H * W * c1, (int64_t)W * c1, (int64_t)c1, H, W, c1};
    const int ctot = c0 + (a1 ? c1 : 0);
    const int r = ksize / 2;
    int np = 0;
    for (int ky = 0; ky < ksize; ++ky)
      for (int kx = 0; kx < ksize; ++kx) {
        const int tapi = ky * ksize + kx;
        d.phases[np++] = {0, ky - r, kx - r, tapi * ctot};
        if (a1) d.phases[np++] = {1, ky - r, kx - r, tapi * ctot + c0};
      }
    d.n_phase = np;
    d.weight = w; d.Ktot = (int64_t)ksize * ksize * ctot;
    d.bias = bias; d.row_scale = row_scale; d.residual = residual; d.act = act; d.out_mode = out_mode; d.out = out;
    d.gn_partials = (conv_impl & 3) ? nullptr : partials;
    d.splitk_ws = sk_ws; d.splitk_ws_bytes = sk_bytes;
    run((conv_impl & 1) ? srgd_conv_direct(&d, st) : srgd_conv_igemm(&d, st));
  }

  // conv3x3 whose epilogue leaves the GroupNorm partial records in `part` (debug paths: statistics in `stats`)
  void conv_gn(const void* a0, int c0, const void* a1, int c1, int H, int W, const void* w, const float* bias, int cout,
               void* out, float* part, float* stats) {
    const bool fused = fused_stats(H, W);
    conv(a0, c0, a1, c1, H, W, 3, w, bias, cout, out, fused ? part : nullptr, nullptr, nullptr, 0, SRGD_OUT_BF16_NHWC);
    if (!dry && ok() && !fused) run(srgd_groupnorm_stats(out, stats, B, H, W, cout, st));
  }
  // The conv epilogue's partial records describe at most 4 samples per 128-pixel tile: feature maps below 32 pixels
  // (inputs under ~48x48 at four levels; the reference accepts anything divisible by 8, model.py:679) take the
  // stand-alone statistics kernel instead.
  bool fused_stats(int H, int W) const { return !(conv_impl & 3) && H * W >= 32; }

  // A conv whose second source has only Bb < B rows (the same rows for every group of Bb samples: init_conv's output
  // shared by the two halves of a class-guidance batch): B / Bb launches over Bb samples each.  The GroupNorm partial
  // records of the groups are laid out as one B-row launch would write them (one record set per sample; H*W >= 128).
  void conv_bcast(const void* a0, int c0, const void* a1, int c1, int Bb, int H, int W, int ksize, const void* w,
                  const float* bias, int cout, void* out, float* part) {
    const size_t px = (size_t)Bb * H * W;
    const size_t rec = (size_t)srgd_conv_m_tiles(Bb, H, W) * 16;
    for (int g = 0; g < B / Bb; ++g)
      conv(reinterpret_cast<const uint8_t*>(a0) + g * px * c0 * 2, c0, a1, c1, H, W, ksize, w, bias, cout,
           reinterpret_cast<uint8_t*>(out) + g * px * cout * 2, part ? part + g * rec : nullptr, nullptr, nullptr, 0,
           SRGD_OUT_BF16_NHWC, Bb);
  }

  // ResnetBlock (model.py:261-285).  Consumes nothing; returns a fresh [B][H][W][cout] buffer.
  // inv_out (optional): receives 1/||row|| of the block output for the attention block that follows.
  // eps_out (optional, last block of the network only): the final 1x1 conv is fused into the second GroupNorm pass
  // (srgd_groupnorm_apply_final); the block output is then not materialised and nullptr is returned.
  // GroupNorm statistics never get a launch of their own on the product path: the conv epilogue writes one
  // 64-byte partial record per 128-pixel tile and every block of the apply kernel folds its sample's records.
  // Class-guidance sharing (Ba / Bb, both default to "all B rows"):
  //   Ba < B: xa has only Ba rows, the same for every group of Ba samples (the first block of the network on the
  //           shared init_conv output): conv1 and its GroupNorm statistics are computed once per group member, the
  //           apply pass broadcasts them to the B rows (whose scale / shift differ), the identity residual is
  //           broadcast too.  Needs a single source and no res_conv.
  //   Bb < B: xb has only Bb rows (the final block's concat with init_conv's output): its consumers run as B / Bb
  //           launches (conv_bcast).
  void* resblock(const ResP& r, const void* xa, const void* xb, int H, int W, float* inv_out = nullptr,
                 float* eps_out = nullptr, int Ba = -1, int Bb = -1) {
    const size_t M = (size_t)B * H * W;
    const bool fused = fused_stats(H, W);
    const bool shared_in = Ba > 0 && Ba < B, bcast_b = Bb > 0 && Bb < B;
    void* c1 = alloc(M * r.cout * 2);
    float* stats = reinterpret_cast<float*>(alloc((size_t)B * 8 * 2 * sizeof(float)));
    float* part = reinterpret_cast<float*>(alloc((size_t)srgd_conv_m_tiles(B, H, W) * 8 * 2 * sizeof(float)));
    const float* st_arg = fused ? nullptr : stats;
    const float* pt_arg = fused ? part : nullptr;
    if (shared_in) {
      void* c1s = alloc((size_t)Ba * H * W * r.cout * 2);
      conv(xa, r.cin0, nullptr, 0, H, W, 3, r.c1_w, r.c1_b, r.cout, c1s, part, nullptr, nullptr, 0, SRGD_OUT_BF16_NHWC,
           Ba);
      if (!dry && ok())
        run(srgd_groupnorm_apply(c1s, Ba, nullptr, part, r.n1_g, r.n1_b, ss + r.ss_off, u.ss_total, nullptr, c1,
                                 nullptr, B, H, W, r.cout, st));
      ar.release(c1s);
    } else {
      if (bcast_b) conv_bcast(xa, r.cin0, xb, r.cin1, Bb, H, W, 3, r.c1_w, r.c1_b, r.cout, c1, part);
      else conv_gn(xa, r.cin0, xb, r.cin1, H, W, r.c1_w, r.c1_b, r.cout, c1, part, stats);
      if (!dry && ok())
        run(srgd_groupnorm_apply(c1, B, st_arg, pt_arg, r.n1_g, r.n1_b, ss + r.ss_off, u.ss_total, nullptr, c1, nullptr,
                                 B, H, W, r.cout, st));
    }
    void* c2 = alloc(M * r.cout * 2);
    conv_gn(c1, r.cout, nullptr, 0, H, W, r.c2_w, r.c2_b, r.cout, c2, part, stats);
    ar.release(c1);
    const void* resid = xa;
    void* rbuf = nullptr;
    if (r.res_w != nullptr) {                                           // res_conv 1x1 (model.py:271)
      rbuf = alloc(M * r.cout * 2);
      if (bcast_b) conv_bcast(xa, r.cin0, xb, r.cin1, Bb, H, W, 1, r.res_w, r.res_b, r.cout, rbuf, nullptr);
      else
        conv(xa, r.cin0, xb, r.cin1, H, W, 1, r.res_w, r.res_b, r.cout, rbuf, nullptr, nullptr, nullptr, 0,
             SRGD_OUT_BF16_NHWC);
      resid = rbuf;
    }
    if (eps_out != nullptr) {
      if (!dry && ok())
        run(srgd_groupnorm_apply_final(c2, st_arg, pt_arg, r.n2_g, r.n2_b, resid, u.final_w, u.final_b, eps_out, B, H,
                                       W, r.cout, st));
      ar.release(rbuf);
      ar.release(part);
      ar.release(stats);
      ar.release(c2);
      return nullptr;
    }
    if (!dry && ok())
      run(srgd_groupnorm_apply_ex(c2, B, st_arg, pt_arg, r.n2_g, r.n2_b, nullptr, 0, resid,
                                  (shared_in && rbuf == nullptr) ? Ba : B, c2, inv_out, B, H, W, r.cout, st));
    ar.release(rbuf);
    ar.release(part);
    ar.release(stats);
    tap(r.name, c2, M * r.cout * 2);
    return c2;
  }

  // attn(x) + x  (model.py:703, 709, 718).  Returns a fresh buffer.
  // true if the resblock feeding this attention block should emit the per-pixel 1/||x|| (fused-LA path, C <= 256)
  bool wants_inv(const AttnP& a, int H, int W) const {
    return !a.full && !(conv_impl & 1) && srgd_linear_attention_block_supported(H * W, a.C, u.cfg.heads) &&
           (a.C == 128 || a.C == 256) && (H * W) % 4 == 0;
  }

  void* attention(const AttnP& a, const void* x, int H, int W, const float* inv_in = nullptr) {
    const size_t M = (size_t)B * H * W;
    const int hid = u.hidden;
    if (!a.full && !(conv_impl & 1) && srgd_linear_attention_block_supported(H * W, a.C, u.cfg.heads)) {
      // fused tcgen05 block: q/k/v/o stay on chip (linattn_fused.cu)
      const size_t wsb = srgd_linear_attention_block_workspace(B, H * W, a.C, u.cfg.heads);
      void* lws = alloc(wsb);
      void* out = alloc(M * a.C * 2);
      if (!dry && ok())
        run(srgd_linear_attention_block(x, inv_in, a.qkv_w, a.out_w, a.out_b, a.out_g, out, B, H * W, a.C, u.cfg.heads,
                                        lws, wsb, st));
      ar.release(lws);
      tap(a.name, out, M * a.C * 2);
      return out;
    }
    float* inv = reinterpret_cast<float*>(alloc(M * sizeof(float)));
    if (!dry && ok()) run(srgd_pixel_inv_norm(x, inv, (int64_t)M, a.C, st));
    void* qkv = alloc(M * 3 * hid * 2);
    // RMSNorm folded: g*sqrt(C) lives in qkv_w's columns, 1/||x|| is the per-pixel row scale
    conv(x, a.C, nullptr, 0, H, W, 1, a.qkv_w, nullptr, 3 * hid, qkv, nullptr, inv, nullptr, 0, SRGD_OUT_BF16_NHWC);
    void* ao = alloc(M * hid * 2);
    void* out = nullptr;
    if (a.full) {
      if (!dry && ok()) {
        if (!(conv_impl & 1) && srgd_attention_tc_supported(H * W, u.cfg.heads))
          run(srgd_attention_tc(qkv, ao, B, H * W, u.cfg.heads, st));       // tcgen05 flash attention
        else
          run(srgd_attention(qkv, ao, B, H * W, u.cfg.heads, st));          // CUDA-core path (small N / debug)
      }
      ar.release(qkv);
      ar.release(inv);
      out = alloc(M * a.C * 2);
      conv(ao, hid, nullptr, 0, H, W, 1, a.out_w, a.out_b, a.C, out, nullptr, nullptr, x, 0, SRGD_OUT_BF16_NHWC);
    } else {
      const size_t wsb = srgd_linear_attention_workspace(B, H * W, u.cfg.heads);
      void* lws = alloc(wsb);
      if (!dry && ok()) run(srgd_linear_attention(qkv, ao, B, H * W, u.cfg.heads, lws, wsb, st));
      ar.release(lws);
      ar.release(qkv);
      ar.release(inv);
      out = alloc(M * a.C * 2);
      conv(ao, hid, nullptr, 0, H, W, 1, a.out_w, a.out_b, a.C, out, nullptr, nullptr, nullptr, 0, SRGD_OUT_BF16_NHWC);
      if (!dry && ok()) run(srgd_rmsnorm_residual(out, a.out_g, x, out, (int64_t)M, a.C, st));
    }
    ar.release(ao);
    tap(a.name, out, M * a.C * 2);
    return out;
  }

  // Downsample (model.py:106-110): 2x2 space-to-depth expressed as four strided sources
  void* downsample(const ConvP& c, const void* x, int H, int W) {
    const int cin = c.cin / 4, Ho = H / 2, Wo = W / 2;
    void* out = alloc((size_t)B * Ho * Wo * c.cout * 2);
    if (!dry && ok()) {
      srgd_conv_desc d;
      memset(&d, 0, sizeof(d));
      d.B = B; d.Ho = Ho; d.Wo = Wo; d.Cout = c.cout;
      d.n_src = 4; d.n_phase = 4;
      for (int i = 0; i < 4; ++i) {
        const int p1 = i >> 1, p2 = i & 1;
        d.srcs[i] = {reinterpret_cast<const uint8_t*>(x) + ((size_t)p1 * W + p2) * cin * 2, (int64_t)H * W * cin,
                     (int64_t)2 * W * cin, (int64_t)2 * cin, Ho, Wo, cin};
        d.phases[i] = {i, 0, 0, i * cin};
      }
      d.weight = c.w; d.Ktot = c.cin; d.bias = c.b; d.out = out; d.out_mode = SRGD_OUT_BF16_NHWC;
      d.splitk_ws = sk_ws; d.splitk_ws_bytes = sk_bytes;
      run((conv_impl & 1) ? srgd_conv_direct(&d, st) : srgd_conv_igemm(&d, st));
    }
    tap(c.name, out, (size_t)B * Ho * Wo * c.cout * 2);
    return out;
  }
};

static int forward_impl(srgd_unet& u, Arena& ar, bool dry, const float* x, const float* cond, const float* log_snr,
                        const int32_t* labels, int n_cond_rows, int Bx, float* eps, int B, int H, int W, int conv_impl,
                        srgd_stream_t st) {
  Fwd f{u, ar, dry, conv_impl, st, B};
  const srgd_unet_config& c = u.cfg;
  const int n = c.n_stages;
  const int dim = c.dim, td = u.time_dim;
  const size_t M0 = (size_t)B * H * W;

  // ---- split-K workspace of the conv launches: flags zeroed once per forward (the kernels re-arm them) ----
  f.sk_bytes = (int64_t)srgd_conv_splitk_workspace_bytes();
  f.sk_ws = f.alloc((size_t)f.sk_bytes);
  if (!dry && f.ok()) {
    cudaError_t e = cudaMemsetAsync(f.sk_ws, 0, 4096, as_stream(st));
    if (e != cudaSuccess) f.run(fail_cuda(e, "split-K flag reset"));
  }

  // ---- embeddings (model.py:689-694 and every ResnetBlock.mlp, 264-267/277-279) ----
  const int fdim = c.fixed_sinusoidal ? c.dim : c.sinu_dim + 1;         // model.py:596-601
  float* feats = reinterpret_cast<float*>(f.alloc((size_t)B * fdim * 4));
  float* t1 = reinterpret_cast<float*>(f.alloc((size_t)B * td * 4));
  float* t = reinterpret_cast<float*>(f.alloc((size_t)B * td * 4));
  float* ss = reinterpret_cast<float*>(f.alloc((size_t)B * u.ss_total * 4));
  f.ss = ss;
  if (!dry && f.ok()) {
    if (c.fixed_sinusoidal) f.run(srgd_sinusoidal_pos_emb(log_snr, u.time_freq, feats, B, c.dim / 2, st));
    else f.run(srgd_fourier_features(log_snr, u.time_freq, feats, B, c.sinu_dim / 2, st));
    f.run(srgd_dense_rows(feats, u.time_w1, u.time_b1, t1, B, td, fdim, 0, 0, st));
    f.run(srgd_dense_rows(t1, u.time_w2, u.time_b2, t, B, td, td, 2, 0, st));
    if (labels != nullptr && u.class_table != nullptr)
      f.run(srgd_add_class_rows(t, u.class_table, labels, B, td, c.num_classes, st));
    f.run(srgd_dense_rows(t, u.ss_w, u.ss_b, ss, B, u.ss_total, td, 1, 0, st));
  }
  f.tap("t_emb", t, (size_t)B * td * 4);
  ar.release(feats);
  ar.release(t1);

  // ---- class guidance: rows b and b + Bx read the same x and the same condition (model.py:3151-3154 runs the U-Net
  // twice on them) and differ only in the embedding, which first enters at the first block's scale / shift: the input
  // pack, init_conv and the first conv3x3 + its GroupNorm statistics are computed for Bx rows and broadcast ----
  const char* share_knob = getenv("SRGD_CFG_SHARE");                   // test knob: "0" = compute both halves in full
  const bool share_env = !(share_knob != nullptr && share_knob[0] == '0');
  const bool share = share_env && B == 2 * Bx && (n_cond_rows == B || n_cond_rows == 0) && !(conv_impl & 3) &&
                     tile_geom(1, H, W).tn_log2 == 0 && u.downs[0].r0.res_w == nullptr;
  const int Bs = share ? Bx : B;
  const size_t M0s = (size_t)Bs * H * W;

  // ---- init conv 7x7 (model.py:686): row-im2col pack + 7 vertical taps of 64 channels ----
  void* pk = f.alloc(M0s * 64 * 2);
  if (!dry && f.ok()) f.run(srgd_pack_input(x, cond, share ? (n_cond_rows ? Bs : 0) : n_cond_rows, Bx, pk, Bs, H, W, st));
  void* r = f.alloc(M0s * dim * 2);
  if (!dry && f.ok()) {
    srgd_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.B = Bs; d.Ho = H; d.Wo = W; d.Cout = dim;
    d.n_src = 1; d.n_phase = 7;
    d.srcs[0] = {pk, (int64_t)H * W * 64, (int64_t)W * 64, 64, H, W, 64};
    for (int i = 0; i < 7; ++i) d.phases[i] = {0, i - 3, 0, i * 64};
    d.weight = u.init_w; d.Ktot = 7 * 64; d.bias = u.init_b; d.out = r; d.out_mode = SRGD_OUT_BF16_NHWC;
    f.run((conv_impl & 1) ? srgd_conv_direct(&d, st) : srgd_conv_igemm(&d, st));
  }
  ar.release(pk);
  f.tap("init_conv", r, M0s * dim * 2);

  // ---- down path (model.py:698-706) ----
  std::vector<void*> skips;
  std::vector<int> skip_c;
  void* xcur = r;                       // r stays alive until the final concat
  int h = H, w = W;
  for (int i = 0; i < n; ++i) {
    const Stage& s = u.downs[i];
    void* a = f.resblock(s.r0, xcur, nullptr, h, w, nullptr, nullptr, (i == 0 && share) ? Bs : -1);
    if (xcur != r) ar.release(xcur);
    skips.push_back(a); skip_c.push_back(s.r0.cout);
    float* inv = f.wants_inv(s.attn, h, w) ? reinterpret_cast<float*>(f.alloc((size_t)B * h * w * sizeof(float))) : nullptr;
    void* b2 = f.resblock(s.r1, a, nullptr, h, w, inv);
    void* at = f.attention(s.attn, b2, h, w, inv);
    ar.release(b2);
    ar.release(inv);
    skips.push_back(at); skip_c.push_back(s.attn.C);
    if (!s.last) {
      xcur = f.downsample(s.resample, at, h, w);
      h /= 2; w /= 2;
    } else {
      xcur = f.alloc((size_t)B * h * w * s.resample.cout * 2);
      f.conv(at, s.resample.cin, nullptr, 0, h, w, 3, s.resample.w, s.resample.b, s.resample.cout, xcur, nullptr,
             nullptr, nullptr, 0, SRGD_OUT_BF16_NHWC);
      f.tap(s.resample.name, xcur, (size_t)B * h * w * s.resample.cout * 2);
    }
  }
  // ---- middle (model.py:708-710) ----
  {
    void* a = f.resblock(u.mid1, xcur, nullptr, h, w);
    ar.release(xcur);
    void* at = f.attention(u.mid_attn, a, h, w);
    ar.release(a);
    xcur = f.resblock(u.mid2, at, nullptr, h, w);
    ar.release(at);
  }
  // ---- up path (model.py:712-720) ----
  for (int i = 0; i < n; ++i) {
    const Stage& s = u.ups[i];
    void* sk = skips.back(); skips.pop_back(); skip_c.pop_back();
    void* a = f.resblock(s.r0, xcur, sk, h, w);
    ar.release(xcur); ar.release(sk);
    sk = skips.back(); skips.pop_back(); skip_c.pop_back();
    float* inv = f.wants_inv(s.attn, h, w) ? reinterpret_cast<float*>(f.alloc((size_t)B * h * w * sizeof(float))) : nullptr;
    void* b2 = f.resblock(s.r1, a, sk, h, w, inv);
    ar.release(a); ar.release(sk);
    void* at = f.attention(s.attn, b2, h, w, inv);
    ar.release(b2);
    ar.release(inv);
    if (!s.last) {                                                      // PixelShuffleUpsample (model.py:70-98)
      const int cq = s.resample.cout / 4;
      xcur = f.alloc((size_t)B * (2 * h) * (2 * w) * cq * 2);
      f.conv(at, s.resample.cin, nullptr, 0, h, w, 1, s.resample.w, s.resample.b, s.resample.cout, xcur, nullptr,
             nullptr, nullptr, 1, SRGD_OUT_PIXEL_SHUFFLE);
      h *= 2; w *= 2;
      f.tap(s.resample.name, xcur, (size_t)B * h * w * cq * 2);
    } else {
      xcur = f.alloc((size_t)B * h * w * s.resample.cout * 2);
      f.conv(at, s.resample.cin, nullptr, 0, h, w, 3, s.resample.w, s.resample.b, s.resample.cout, xcur, nullptr,
             nullptr, nullptr, 0, SRGD_OUT_BF16_NHWC);
      f.tap(s.resample.name, xcur, (size_t)B * h * w * s.resample.cout * 2);
    }
    ar.release(at);
  }
  // ---- head (model.py:722-725) ----
  const bool fuse_final = !(conv_impl & 1) && u.taps.empty() && dim == 128 && c.channels == 3 && (h * w) % 2 == 0;
  void* fr = f.resblock(u.final_res, xcur, r, h, w, nullptr, fuse_final ? (dry ? reinterpret_cast<float*>(1) : eps) : nullptr,
                        -1, share ? Bs : -1);
  ar.release(xcur);
  ar.release(r);
  if (!fuse_final) {
    if (!dry && f.ok()) f.run(srgd_final_conv(fr, u.final_w, u.final_b, eps, B, h, w, dim, c.channels, st));
    ar.release(fr);
  }
  ar.release(t);
  ar.release(ss);
  ar.release(f.sk_ws);
  return f.rc;
}

static int check_shape(const srgd_unet* u, int B, int H, int W) {
  SRGD_REQUIRE(u != nullptr, "unet: null handle");
  const int factor = 1 << (u->cfg.n_stages - 1);
  SRGD_REQUIRE(B > 0 && H > 0 && W > 0, "unet: bad shape");
  SRGD_REQUIRE(H % factor == 0 && W % factor == 0,
               "your input dimensions (%d, %d) need to be divisible by %d, given the unet", H, W, factor);
  return SRGD_OK;
}

}  // namespace srgd

using namespace srgd;

extern "C" int srgd_unet_param_count(const srgd_unet_config* cfg) {
  if (validate_cfg(cfg) != SRGD_OK) return SRGD_E_ARG;
  srgd_unet tmp;
  tmp.cfg = *cfg;
  walk(tmp, false);
  return (int)tmp.names.size();
}

extern "C" const char* srgd_unet_param_name(const srgd_unet_config* cfg, int index) {
  static thread_local std::string s;
  if (validate_cfg(cfg) != SRGD_OK) return nullptr;
  srgd_unet tmp;
  tmp.cfg = *cfg;
  walk(tmp, false);
  if (index < 0 || index >= (int)tmp.names.size()) return nullptr;
  s = tmp.names[index];
  return s.c_str();
}

extern "C" int srgd_unet_create(const srgd_unet_config* cfg, const void* const* params_dev, int n_params,
                                srgd_unet** out) {
  int rc = validate_cfg(cfg);
  if (rc) return rc;
  SRGD_REQUIRE(params_dev && out, "unet_create: null argument");
  srgd_unet* u = new srgd_unet();
  u->cfg = *cfg;
  walk(*u, false);
  if ((int)u->names.size() != n_params) {
    set_error("unet_create: expected %d parameter pointers, got %d", (int)u->names.size(), n_params);
    delete u;
    return SRGD_E_ARG;
  }
  for (int i = 0; i < n_params; ++i) {
    if (params_dev[i] == nullptr || ((uintptr_t)params_dev[i] % 16) != 0) {
      set_error("unet_create: parameter %d (%s) is null or not 16-byte aligned", i, u->names[i].c_str());
      delete u;
      return SRGD_E_ARG;
    }
    u->ptrs.push_back(params_dev[i]);
  }
  walk(*u, true);
  *out = u;
  return SRGD_OK;
}

extern "C" void srgd_unet_destroy(srgd_unet* u) { delete u; }

extern "C" size_t srgd_unet_workspace_bytes(const srgd_unet* u, int32_t B, int32_t H, int32_t W) {
  if (check_shape(u, B, H, W) != SRGD_OK) return 0;
  // the product plan and the debug plan (conv_impl bit 0: unfused attention) differ: size for both
  size_t need = 0;
  for (int impl = 0; impl < 2; ++impl)
    for (int shared = 0; shared < ((B % 2 == 0) ? 2 : 1); ++shared) {       // plain plan and class-guidance sharing plan
      Arena ar(nullptr, (size_t)1 << 60);
      forward_impl(*const_cast<srgd_unet*>(u), ar, true, nullptr, nullptr, nullptr, nullptr, shared ? B : 0,
                   shared ? B / 2 : B, nullptr, B, H, W, impl, nullptr);
      if (ar.high_water() > need) need = ar.high_water();
    }
  return need + 256;
}

extern "C" int srgd_unet_forward(srgd_unet* u, const float* x_dev, const float* cond_dev, const float* log_snr_dev,
                                 const int32_t* labels_dev, int32_t n_cond_rows, int32_t Bx, float* eps_dev, int32_t B,
                                 int32_t H, int32_t W, void* workspace_dev, size_t workspace_bytes, int32_t conv_impl,
                                 srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  rc = check_shape(u, B, H, W);
  if (rc) return rc;
  SRGD_REQUIRE(x_dev && log_snr_dev && eps_dev && workspace_dev, "unet_forward: null argument");
  SRGD_REQUIRE(Bx > 0 && Bx <= B && B % Bx == 0, "unet_forward: B=%d must be a multiple of Bx=%d", B, Bx);
  SRGD_REQUIRE(((uintptr_t)workspace_dev % 256) == 0, "unet_forward: workspace must be 256-byte aligned");
  Arena ar(workspace_dev, workspace_bytes);
  const long before = g_launches;
  rc = forward_impl(*u, ar, false, x_dev, cond_dev, log_snr_dev, labels_dev, n_cond_rows, Bx, eps_dev, B, H, W,
                    conv_impl, stream);
  u->last_launches = g_launches - before;
  return rc;
}

extern "C" int srgd_unet_set_tap(srgd_unet* u, const char* name, void* out_dev, size_t out_bytes) {
  SRGD_REQUIRE(u != nullptr, "unet_set_tap: null handle");
  if (name == nullptr) {
    u->taps.clear();
    return SRGD_OK;
  }
  SRGD_REQUIRE(out_dev != nullptr && out_bytes > 0, "unet_set_tap: null buffer");
  u->taps.push_back({std::string(name), out_dev, out_bytes});
  return SRGD_OK;
}

extern "C" int srgd_unet_last_launch_count(const srgd_unet* u) { return u ? (int)u->last_launches : 0; }
