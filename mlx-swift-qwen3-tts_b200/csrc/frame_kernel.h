// Persistent frame kernel ("megakernel") for the batch-1..NS decode loop: ONE cooperative launch runs n frames of
// Model/Qwen3Talker.swift:464-562 — code0 sample, 15 code-predictor passes, frame finalize, talker step — with every
// Qwen3DecoderLayer phase handing its results to the next through tagged values in L2 instead of a kernel boundary.  Weights are streamed by a producer
// warp (1-D TMA bulk copies into a shared-memory ring, several phases ahead of the math), so the dependency chain only
// ever waits on L2-resident activations.
#pragma once
#include "kernels.h"

namespace q3 {

struct MegaLinear {       // one QuantizedLayerFactory.linear leaf in execution order
  const void* w;          // packed (MLX affine) or dense rows
  const void* scales;     // quantised only: [out][in/group]
  const void* biases;
  const float* bias;      // optional Linear bias, fp32 [out]
  int out_eff;            // rows of one sub-matrix (SwiGLU: out / 2)
  int nsub;               // 1, or 2 for [gate ; up]
  int in;                 // K
  int row_bytes;          // bytes of one weight row
  int srow_bytes;         // bytes of one row of scales (== biases); 0 for dense formats
  int unit;               // row granularity of a CTA's slice (keeps every bulk copy 16-byte aligned)
  int sdt;                // dtype of scales / biases
  int group;
  int ubase, urem;        // slice of CTA c: ubase + (c < urem) units starting at unit c*ubase + min(c, urem)
  int rch;                // rows per ring chunk
  int group_shift;        // log2(group)
  const float* norm_w;    // RMSNorm weight fused into this linear's prologue (qkv, gate|up, heads), or null
  const float* q_norm;    // qkv entries: per-head norms of the attention phase that follows
  const float* k_norm;
  // ---- the phase this linear runs in, precomputed so the kernel's loop top is a handful of shared-memory reads
  int in_kind;            // InKind of the staged rows
  int pass;               // code-predictor pass / unit index 0..15 (15 = talker); rows per slot for the head's IN_GX_LAST
  int epi;                // EpiKind
  int out_sel;            // 0 residual x, 1 qkv, 2 SwiGLU activations, 3 logits
  int flags;              // MegaFlags
  int layer;
  int uidx;               // 0..14 code-predictor pass, 15 talker step
  int tkind;              // trace label: 0 mtp, 1 qkv, 2 o, 3 gate|up, 4 down, 5 head
  int pad_[2];            // sizeof == 144: the kernel moves descriptors as 16-byte vectors
};
enum MegaFlags {
  MF_UNIT_START = 1,      // first linear of a unit: the code of this unit is sampled before it
  MF_ATTN = 2,            // qkv: an attention phase follows
  MF_TALKER = 4,
  MF_HEAD = 8,            // lm_head / codec_head: rows = one per slot
  MF_KEEP_RAW = 16,       // staged rows are the residual of the following o / down projection
  MF_ROWS2 = 32,          // two rows per slot (code-predictor pass 0)
  MF_FINALIZE = 64        // talker layer 0: CTA 0 records the finished frame before the attention phase
};
static_assert(sizeof(MegaLinear) % 16 == 0, "MegaLinear must be a whole number of 16-byte vectors");

struct MegaStack {
  int hidden, layers, heads, kv_heads, inter;
  float eps;
  const float* final_norm;
  const float* inv_freq;
  float* k;
  float* v;
  size_t slot_stride, layer_stride;
  int capacity;
  int nsplit;               // key splits per (slot, kv head) attention item
};

struct MegaParams {
  const MegaLinear* lin;    // device array, one frame's linears in execution order
  int n_lin;
  MegaStack cp, tk;
  int has_mtp, H, Hcp, V, Vc;
  Embedding codec;
  const Embedding* cp_emb;  // device array [15]
  SlotState* st;
  int n_slots;
  int* cur_codes;
  int* frames;
  const int* forced;
  int max_frames;
  unsigned* sets;
  int set_words;
  const float* trailing;
  int max_trailing;
  const float* tts_pad;
  float *hlast, *logits0, *cplogits, *dump0, *dumpcp;
  // LL exchange buffers: 8-byte (value, tag) elements, zeroed per launch (ex_base, ex_bytes)
  unsigned long long *ex_x, *ex_qkv, *ex_part, *ex_act, *ex_logit, *ex_msg;
  int ld_x, ld_qkv, ld_act, ld_logit;
  int part_stride;          // elements per (row, split) partial: heads*128 + 2*heads, padded to 4
  void* ex_base;
  size_t ex_bytes;
  int n_frames, window, eos_id, pad_id;
  // shared-memory plan (bytes from the 1024-aligned base)
  int slot_bytes, n_ring;
  int off_xs, off_xsum, off_xraw, off_red, off_bar, off_dsc, off_hl, off_rope, off_part;
  int xh_stride;            // halfs per activation column of the tensor-core path (K max + 32 padding)
  int raw_ld;               // floats per row of the raw residual copy
  long long* trace;         // diagnostics: [2 CTAs][trace_stride] cycle stamps, 8 per phase (null = off)
  int trace_stride;
};

constexpr int kMegaMaxSlots = 2;   // utterances per launch on the tensor-core path: 2 x 2 rows, hi + lo fp16 halves = the 8 columns of one MMA n-tile

// false when this checkpoint / option set is outside the kernel's envelope (the CUDA-graph path is used instead)
struct MegaPlan {
  bool ok = false;
  MegaParams p{};
  int fmt = 0;       // WFmt
  int G_cp = 0, G_tk = 0;
  size_t smem = 0;
  int grid = 0;
  int max_slots = 1;  // utterances one launch can serve
};

void init_mega_kernels();
void launch_frame_megakernel(const LaunchCtx& c, const MegaPlan& plan, int n_slots, int n_frames, float* dump0, float* dumpcp);

}  // namespace q3
