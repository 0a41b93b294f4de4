// generator.cu -- native (C++) orchestration of the whole SG-NN generator forward (reference torch/model.py:371-416)
// on top of the kernels of this library.  The reference drives ~130 scn ops + Python glue per pass from Python
// with CPU-side metadata; here one C-ABI call enqueues the ~190 kernels of a pass on one stream and touches the
// host only to read the data-dependent row counts (4 bytes each: coarse site counts, kept-candidate counts).
// Memory comes from a caller-provided bump arena (no cudaMalloc on the path); temporaries are released in
// stream order.  The arithmetic is the fused composition of sgnn_b200/fused.py -- bit-identical to the
// module-by-module path.
#include <stdio.h>
#include <string.h>
#include "common.cuh"

// Row thresholds of the tensor-core paths under SGNN_GEN_TC32 (SgnnGeneratorW::tc32_min_rows / ur_min_rows override them):
//  * site sets with at least UR rows get a unique-row tile plan (conv_ur.cu / conv_urc.cu) for their K = 27, Cout = 16
//    convolutions and their child-mode convolution;
//  * other Cout = 16 convolutions (stride-2, or no plan) with at least TC32 output rows run on the round-1 tcgen05 kernels; below
//    that a launch cannot fill the chip with 128-row tiles and the FFMA kernel's shorter per-tile latency wins.
#define SGNN_DEFAULT_TC32_MIN_ROWS 60000
#define SGNN_DEFAULT_UR_MIN_ROWS 1000
// levels of at most this many rows build their coarse site sets on the side stream (SgnnGeneratorW::overlap_max_rows)
#define SGNN_DEFAULT_OVERLAP_MAX_ROWS 200000

namespace {

struct Arena {
  char* base;
  size_t cap, off, high;
  bool oom;
  void* get(size_t bytes) {
    size_t a = (off + 255) & ~(size_t)255;
    size_t end = a + (bytes ? bytes : 1);
    if (end > high) high = end;
    if (end > cap) { oom = true; return nullptr; }
    off = end;
    return base + a;
  }
};

struct Ctx {
  Arena ar;
  cudaStream_t st;        // the stream helper calls enqueue on: main_st, or side_st between use_side() and use_main()
  void* stream;
  cudaStream_t main_st, side_st;   // side_st == nullptr: no overlap, everything on main_st
  bool side_on;                    // between fork_side() and join_side() of a level small enough to overlap
  long long overlap_max_rows;
  bool profile;
  bool tc32;
  bool compact;    // compact rulebook + conv_sp.cu on the encoder's input level
  bool phases;
  int n_ph;
  int n_ev;
  const SgnnGeneratorW* w;
  long long tc32_min_rows, ur_min_rows;
};

// event pool for SGNN_GEN_PROFILE (pairs around every sgnn_conv_forward of the pass)
static thread_local cudaEvent_t g_ev[512];
static thread_local int g_ev_made = 0;
// shapes + durations of the convolutions of the last profiled pass (sgnn_generator_profile_entry)
struct ConvRec { int64_t n_out; int cin, cout, K, child, tc; float ms; };
static thread_local ConvRec g_rec[256];
static thread_local int g_nrec = 0;

// SGNN_GEN_PHASES: one CUDA event at every phase boundary of the pass (~45 marks); the time between consecutive marks is what
// the phase really costs in situ -- kernels, launch gaps, memsets and host-read stalls included (sgnn_generator_phase_entry).
struct PhaseRec { char name[48]; float ms; };
static thread_local cudaEvent_t g_ph_ev[96];
static thread_local int g_ph_made = 0;
static thread_local PhaseRec g_ph[96];
static thread_local int g_nph = 0;

struct Epi {
  float* out; int ld; const float* scale; const float* shift; int relu;
};
static Epi epi(float* out, int ld, const float* scale = nullptr, const float* shift = nullptr, int relu = 0) {
  Epi e; e.out = out; e.ld = ld; e.scale = scale; e.shift = shift; e.relu = relu; return e;
}
static Epi epi_bn(float* out, int ld, const SgnnBnFold& bn, int off = 0) {
  return epi(out, ld, bn.scale + off, bn.shift + off, 1);
}
static const Epi kNoEpi = {nullptr, 0, nullptr, nullptr, 0};

struct Level {
  SgnnGrid g;
  int32_t* coords;
  int32_t* nbr;
  void* plan;      // unique-row tile plan of nbr (conv_ur.cu) or nullptr
  int32_t* slots;  // compact rulebook (grid.cu COMPACT / conv_sp.cu) instead of nbr, or nullptr
  uint8_t* cnt;
  bool unique;     // no two rows share a coordinate (every set except the caller's input)
  int64_t n;
  int dims[3];
};

#define RC(call)            \
  do {                      \
    int rc__ = (call);      \
    if (rc__) return rc__;  \
  } while (0)
#define ALLOC(var, type, count)                                   \
  type* var = (type*)c.ar.get((size_t)(count) * sizeof(type));    \
  if (!var) return SGNN_E_NOMEM

// pinned landing zone for the 4-byte count reads (pageable destinations are staged by the driver)
static thread_local int32_t* g_pinned = nullptr;   // per host thread: concurrent forwards on different streams
static int pinned_slots(int32_t** p) {
  if (!g_pinned) SGNN_CUDA(cudaHostAlloc((void**)&g_pinned, 64, cudaHostAllocDefault));
  *p = g_pinned;
  return SGNN_OK;
}

static int read_i32(Ctx& c, const int32_t* dev, int32_t* host) {
  int32_t* pin;
  RC(pinned_slots(&pin));
  SGNN_CUDA(cudaMemcpyAsync(pin, dev, 4, cudaMemcpyDeviceToHost, c.st));
  SGNN_CUDA(cudaStreamSynchronize(c.st));
  *host = pin[0];
  return SGNN_OK;
}

// ---- side stream for the SMALL levels.  The coarse site sets of a level (masks, ranks, coordinates, strided rulebooks,
// neighbour tables, plans: ~20 tiny launches) depend only on the level's own grid, not on its features.  On levels of a few
// thousand rows every launch is pure latency (~5 us each, profiles/r02_phases_*.txt), so there they are built on a second stream
// while the level's first convolutions run on the caller's stream.  Only there: on the large levels the small grid kernels would
// share SMs with the persistent one-CTA-per-SM tensor-core convolutions and delay their CTAs (measured: 4.44 ms per step
// single-stream, 4.64 / 4.71 ms with the levels of >= 100 k / >= 300 k rows overlapped, profiles/r02_ab_runs.txt).
// fork: side waits for main's tail; join: main waits for side's tail.  Scratch handed out by the arena is never recycled within
// a pass (the bump pointer only grows), so memory used on one stream is never re-issued to the other.
static thread_local cudaStream_t g_side_stream[64] = {};
static thread_local cudaEvent_t g_fj_ev[16] = {};
static thread_local int g_fj_next = 0;
static cudaStream_t side_stream() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!g_side_stream[dev] && cudaStreamCreateWithFlags(&g_side_stream[dev], cudaStreamNonBlocking) != cudaSuccess) {
    g_side_stream[dev] = nullptr;
    cudaGetLastError();
  }
  return g_side_stream[dev];
}
static int order_after(cudaStream_t waiter, cudaStream_t signaller) {
  cudaEvent_t& e = g_fj_ev[g_fj_next];
  g_fj_next = (g_fj_next + 1) & 15;
  if (!e) SGNN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  SGNN_CUDA(cudaEventRecord(e, signaller));
  SGNN_CUDA(cudaStreamWaitEvent(waiter, e, 0));
  return SGNN_OK;
}
static void use_main(Ctx& c) { c.st = c.main_st; c.stream = (void*)c.main_st; }
static void use_side(Ctx& c) { if (c.side_on) { c.st = c.side_st; c.stream = (void*)c.side_st; } }
static int fork_side(Ctx& c, int64_t level_rows) {     // side work from here on sees everything enqueued on main so far
  c.side_on = c.side_st && level_rows <= c.overlap_max_rows;
  if (!c.side_on) return SGNN_OK;
  RC(order_after(c.side_st, c.main_st));
  use_side(c);
  return SGNN_OK;
}
static int join_side(Ctx& c) {     // main work from here on sees everything enqueued on side so far
  use_main(c);
  if (!c.side_on) return SGNN_OK;
  c.side_on = false;
  return order_after(c.main_st, c.side_st);
}

// phase boundary: everything enqueued since the previous mark is accounted to `name`
static int mark(Ctx& c, const char* name, int level = -1) {
  if (!c.phases || c.n_ph >= 96) return SGNN_OK;
  while (g_ph_made <= c.n_ph) SGNN_CUDA(cudaEventCreate(&g_ph_ev[g_ph_made++]));
  SGNN_CUDA(cudaEventRecord(g_ph_ev[c.n_ph], c.st));
  PhaseRec& r = g_ph[c.n_ph];
  if (level >= 0) snprintf(r.name, sizeof(r.name), "L%d %s", level, name);
  else snprintf(r.name, sizeof(r.name), "%s", name);
  r.ms = 0.f;
  ++c.n_ph;
  return SGNN_OK;
}

static void grid_shape(SgnnGrid* g, int nb, const int dims[3]) {
  memset(g, 0, sizeof(*g));
  g->nb = nb; g->d0 = dims[0]; g->d1 = dims[1]; g->d2 = dims[2];
  g->wx = (dims[2] + 63) / 64;
  g->n_words = (int64_t)nb * dims[0] * dims[1] * g->wx;
}

// does a site set of n rows feeding Cout = cout convolutions get a unique-row tile plan?  (The kernel also takes Cout 8 / 12,
// but on the sparse encoder levels its per-tile overhead loses to the FFMA kernels: 87-95 us against 68-72 us for 420 k rows.)
// (Measured again in round 2 for the 352 k-row Cout = 12 encoder level: no difference, 4.19-4.27 against 4.22-4.24 ms per step.)
static bool wants_plan(const Ctx& c, int64_t n, int cout) { return c.tc32 && cout == 16 && n >= c.ur_min_rows; }

// neighbour table of a site set and, when it qualifies, its tile plan -- in one kernel (sgnn_rulebook_submanifold_plan)
static int rulebook_and_plan(Ctx& c, Level* L, int32_t* nbr, int plan_cout) {
  L->nbr = nbr;
  L->plan = nullptr;
  if (L->n > 0 && wants_plan(c, L->n, plan_cout)) {
    const size_t pb = sgnn_tile_plan_bytes(L->n);
    void* plan = c.ar.get(pb);
    if (!plan) return SGNN_E_NOMEM;
    L->plan = plan;
    return sgnn_rulebook_submanifold_plan(&L->g, L->coords, L->n, nbr, plan, pb, c.stream);
  }
  return sgnn_rulebook_submanifold(&L->g, L->coords, L->n, nbr, c.stream);
}

// a1 + a2: site set from explicit coordinates, with its 27-neighbour table
static int build_level(Ctx& c, const void* coords, int is64, int64_t n, int nb, const int dims[3], int32_t* status,
                       Level* L, bool compact = false, int plan_cout = 0) {
  grid_shape(&L->g, nb, dims);
  for (int i = 0; i < 3; ++i) L->dims[i] = dims[i];
  L->n = n;
  L->unique = status == nullptr;   // internal levels (children of distinct kept sites); the caller's input is checked, not trusted
  ALLOC(mask, uint64_t, L->g.n_words);
  ALLOC(prefix, int32_t, L->g.n_words + 1);
  ALLOC(ror, int32_t, n);
  ALLOC(ci32, int32_t, n * 4);
  L->g.mask = mask; L->g.prefix = prefix; L->g.row_of_rank = ror; L->coords = ci32;
  const size_t sb = sgnn_scan_scratch_bytes(L->g.n_words);
  ALLOC(scr, char, sb);
  RC(sgnn_grid_build(&L->g, coords, is64, n, ci32, status, scr, sb, c.stream));
  ALLOC(nbr, int32_t, 27 * n);
  L->plan = nullptr; L->slots = nullptr; L->cnt = nullptr;
  if (compact && n < (1LL << 27)) {
    ALLOC(cnt, uint8_t, n);
    L->nbr = nullptr; L->slots = nbr; L->cnt = cnt;
    return sgnn_rulebook_submanifold_compact(&L->g, ci32, n, nbr, cnt, c.stream);
  }
  return rulebook_and_plan(c, L, nbr, plan_cout);
}

// a4: stride-2 coarse set of `f` (raster rows) + strided rulebook (+ its own 27-neighbour table), in two phases so
// that a whole pyramid of coarse sets costs ONE host read: begin() enqueues mask + rank scan (depends only on the
// finer mask), the caller reads all counts with a single synchronisation, finish() allocates exactly and enqueues
// coordinate enumeration and rulebooks.
static int coarsen_begin(Ctx& c, const Level& f, Level* L) {
  int dims[3];
  for (int i = 0; i < 3; ++i) dims[i] = f.dims[i] >= 2 ? (f.dims[i] - 2) / 2 + 1 : 0;  // scn output size
  grid_shape(&L->g, f.g.nb, dims);
  for (int i = 0; i < 3; ++i) L->dims[i] = dims[i];
  ALLOC(mask, uint64_t, L->g.n_words);
  ALLOC(prefix, int32_t, L->g.n_words + 1);
  L->g.mask = mask; L->g.prefix = prefix; L->g.row_of_rank = nullptr;
  L->n = -1; L->coords = nullptr; L->nbr = nullptr; L->plan = nullptr; L->slots = nullptr; L->cnt = nullptr;
  L->unique = true;   // one row per set bit
  const size_t sb = sgnn_scan_scratch_bytes(L->g.n_words);
  ALLOC(scr, char, sb);
  RC(sgnn_grid_coarsen(&f.g, &L->g, scr, sb, c.stream));
  return SGNN_OK;
}

// one host read for up to 4 pending levels, in two halves: post() enqueues the 4-byte copies and records an event,
// wait() blocks on THAT EVENT only -- kernels enqueued in between (convolutions that do not depend on the counts) keep the
// GPU busy while the host reads the counts and enqueues the dependent work behind them, so the read costs no GPU idle time.
static thread_local cudaEvent_t g_count_ev = nullptr;
static int counts_post(Ctx& c, Level** lv, int n) {
  int32_t* pin;
  RC(pinned_slots(&pin));
  if (!g_count_ev) SGNN_CUDA(cudaEventCreateWithFlags(&g_count_ev, cudaEventDisableTiming));
  for (int i = 0; i < n; ++i)
    SGNN_CUDA(cudaMemcpyAsync(pin + 1 + i, lv[i]->g.prefix + lv[i]->g.n_words, 4, cudaMemcpyDeviceToHost, c.st));
  SGNN_CUDA(cudaEventRecord(g_count_ev, c.st));
  return SGNN_OK;
}
static int counts_wait(Ctx& c, Level** lv, int n) {
  int32_t* pin;
  RC(pinned_slots(&pin));
  SGNN_CUDA(cudaEventSynchronize(g_count_ev));
  for (int i = 0; i < n; ++i) lv[i]->n = pin[1 + i];
  return SGNN_OK;
}

static int coarsen_finish(Ctx& c, const Level& f, Level* L, int32_t** parent, int32_t** children, bool want_nbr,
                          int plan_cout = 0) {
  const int64_t cnt = L->n;
  ALLOC(cc, int32_t, cnt * 4);
  L->coords = cc;
  ALLOC(par, int32_t, f.n);
  ALLOC(chi, int32_t, cnt * 8);
  *parent = par; *children = chi;
  if (f.unique) {   // coordinates + strided rulebook in one kernel, from the coarse side
    RC(sgnn_grid_coarse_build(&f.g, &L->g, f.n, cnt, cc, par, chi, c.stream));
  } else {          // the caller's input set may hold duplicate coordinates: every duplicate row gets its parent
    if (cnt) RC(sgnn_grid_enumerate(&L->g, cc, c.stream));
    RC(sgnn_rulebook_strided(&L->g, f.coords, f.n, par, chi, cnt, c.stream));
  }
  L->nbr = nullptr;
  if (want_nbr) {
    ALLOC(nbr, int32_t, cnt * 27);
    RC(rulebook_and_plan(c, L, nbr, plan_cout));
  }
  return SGNN_OK;
}

// ---- prepared tensor-core filter banks (one per Cout = 16 convolution weight of the generator, fixed enumeration order)
struct ConvW { const float* w; int K, cin, cout, child; };
static int enumerate_convs(const SgnnGeneratorW* w, ConvW* out) {
  int n = 0;
  auto add = [&](const float* p, int K, int cin, int cout, int child) {
    ConvW c; c.w = p; c.K = K; c.cin = cin; c.cout = cout; c.child = child; out[n++] = c;
  };
  auto res = [&](const SgnnResBlockW& r, int c) { add(r.w0, 27, c, c, 0); add(r.w1, 27, c, c, 0); };
  auto fcnw = [&](const SgnnFcnW& f) {
    for (int b = 0; b < 3; ++b) res(f.blk[b], f.c);
    for (int d = 0; d < 2; ++d) add(f.w_down[d], 8, f.c, f.c, 0);
  };
  for (int l = 0; l < 3; ++l) {
    const SgnnEncLevelW& e = w->enc[l];
    add(e.w_in, 27, e.cin, e.c, 0); res(e.res, e.c); add(e.w_down, 8, e.c, e.c, 0);
  }
  for (int h = 0; h < 3; ++h) {
    const SgnnRefineW& r = w->ref[h];
    add(r.w_in, 27, r.cin, r.c, 0); fcnw(r.fcn); add(r.w_up, 27, 3 * r.c, r.c, 1);
  }
  add(w->surf.w_in, 27, w->surf.cin, w->surf.c, 0); fcnw(w->surf.fcn);
  return n;   // 12 + 3 * 10 + 9 = 51
}
static bool bank_eligible(const ConvW& c) { return c.cout == 16 && c.cin >= 12 && c.cin <= 48 && (!c.child || c.cin == 48); }
// child-mode weights carry two banks: the round-1 kernel's layout, then the unique-row kernel's (conv_urc.cu)
static size_t bank_half(const ConvW& c) { return (sgnn_conv_tc32_workspace_bytes(c.K, c.cin, c.child) + 255) & ~(size_t)255; }
static size_t bank_bytes(const ConvW& c) { return c.child ? 2 * bank_half(c) : bank_half(c); }
// prepared bank of weight `p` inside w->prepared, or nullptr
static void* prepared_bank(const SgnnGeneratorW* w, const float* p, int K, int child) {
  if (!w->prepared) return nullptr;
  ConvW cv[64];
  const int n = enumerate_convs(w, cv);
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    if (!bank_eligible(cv[i])) continue;
    if (cv[i].w == p && cv[i].K == K && cv[i].child == child) return off + bank_bytes(cv[i]) <= w->prepared_bytes ? (char*)w->prepared + off : nullptr;
    off += bank_bytes(cv[i]);
  }
  return nullptr;
}


static int conv(Ctx& c, const float* in, int ld_in, int cin, const int32_t* nbr, int64_t nbr_stride, int K,
                int child, const float* w, int cout, int64_t n_out, const float* res, int ld_res, const Epi& a,
                const Epi& b, int64_t n_in = 0, const void* plan = nullptr, const int32_t* slots = nullptr,
                const uint8_t* cnt = nullptr) {
  SgnnConvArgs x;
  memset(&x, 0, sizeof(x));
  x.in = in; x.ld_in = ld_in; x.dtype = SGNN_F32; x.nbr = nbr; x.nbr_stride = nbr_stride; x.K = K;
  x.child_mode = child; x.weight = w; x.cin = cin; x.cout = cout; x.n_out = n_out; x.residual = res; x.ld_res = ld_res;
  if (n_in == 0 && K == 27 && !child) n_in = n_out;   // submanifold: input rows == output rows
  x.n_in = n_in > 0 && n_in <= 0x7fffffff ? (int32_t)n_in : 0;
  x.a.out = a.out; x.a.ld = a.ld; x.a.relu = a.relu; x.a.scale = a.scale; x.a.shift = a.shift;
  x.b.out = b.out; x.b.ld = b.ld; x.b.relu = b.relu; x.b.scale = b.scale; x.b.shift = b.shift;
  const bool prof = c.profile && c.n_ev + 2 <= 512;   // 256 records
  if (prof) {
    while (g_ev_made < c.n_ev + 2) SGNN_CUDA(cudaEventCreate(&g_ev[g_ev_made++]));
    SGNN_CUDA(cudaEventRecord(g_ev[c.n_ev], c.st));
  }
  int rc = SGNN_E_UNSUPPORTED;
  bool used_tc = false;
  if (slots && cnt) {
    rc = sgnn_conv_forward_compact(&x, slots, cnt, c.stream);
  } else if (c.tc32 && plan && K == 27 && child && cout == 16 && cin == 48) {
    const size_t wb = (size_t)64 * 4608;
    char* bank = (char*)prepared_bank(c.w, w, K, 1);
    void* ws = bank ? bank + ((sgnn_conv_tc32_workspace_bytes(K, cin, 1) + 255) & ~(size_t)255) : c.ar.get(wb);
    if (bank) x.flags |= SGNN_CONV_PREPARED;
    if (!ws) return SGNN_E_NOMEM;
    rc = sgnn_conv_forward_tc32_urc(&x, plan, ws, wb, c.stream);
    used_tc = rc == SGNN_OK;
  } else if (c.tc32 && plan && K == 27 && !child && (cout == 16 || cout == 12 || cout == 8) && cin >= 8 && cin <= 32) {
    const size_t wb = sgnn_conv_tc32_workspace_bytes(K, cin, 0);
    void* ws = prepared_bank(c.w, w, K, 0);
    if (ws) x.flags |= SGNN_CONV_PREPARED;
    else ws = c.ar.get(wb);
    if (!ws) return SGNN_E_NOMEM;
    rc = sgnn_conv_forward_tc32_ur(&x, plan, ws, wb, c.stream);
    used_tc = rc == SGNN_OK;
  } else if (c.tc32 && cout == 16 && cin >= 12 && cin <= 48 && (!child || cin == 48) && n_out >= c.tc32_min_rows) {
    const size_t wb = sgnn_conv_tc32_workspace_bytes(K, cin, child);
    void* ws = prepared_bank(c.w, w, K, child);
    if (ws) x.flags |= SGNN_CONV_PREPARED;
    else ws = c.ar.get(wb);
    if (!ws) return SGNN_E_NOMEM;
    rc = sgnn_conv_forward_tc32(&x, ws, wb, c.stream);
    used_tc = rc == SGNN_OK;
  }
  if (rc == SGNN_E_UNSUPPORTED || rc == SGNN_E_ALIGN) rc = sgnn_conv_forward(&x, c.stream);
  if (prof) {
    SGNN_CUDA(cudaEventRecord(g_ev[c.n_ev + 1], c.st));
    ConvRec& r = g_rec[c.n_ev / 2];
    r.n_out = n_out; r.cin = cin; r.cout = cout; r.K = K; r.child = child; r.tc = used_tc ? 1 : 0; r.ms = 0.f;
    c.n_ev += 2;
  }
  return rc;
}

// y = SMC(BNReLU(SMC(x_bn))) + x_raw   with x_bn = BNReLU_0(x_raw) supplied by the producer of x_raw
static int res_block(Ctx& c, const Level& lv, const SgnnResBlockW& rb, int ch, const float* x_raw, const float* x_bn,
                     const Epi& a, const Epi& b) {
  ALLOC(mid, float, lv.n * ch);
  RC(conv(c, x_bn, ch, ch, lv.nbr, lv.n, 27, 0, rb.w0, ch, lv.n, nullptr, 0, epi_bn(mid, ch, rb.bn1), kNoEpi, 0, lv.plan,
          lv.slots, lv.cnt));
  return conv(c, mid, ch, ch, lv.nbr, lv.n, 27, 0, rb.w1, ch, lv.n, x_raw, ch, a, b, 0, lv.plan, lv.slots, lv.cnt);
}

// FullyConvolutionalNet(reps 1, [c,c,c], residual) + BatchNormReLU(3c): J0 [n, 3c]   (model.py:180-181,255-256)
// The two coarse site sets of a level's FullyConvolutionalNet, first half: masks + rank scans + the count copies, enqueued as
// soon as the level's own grid exists -- BEFORE its first convolution, so that the counts have long arrived when the host asks
// for them (counts_wait, after it has enqueued that convolution and the first residual block) and the host never stalls there.
// On small levels the rest of the construction runs on the side stream (fork_side above).
struct FcnSets { Level lv1, lv2; };
static int fcn_begin(Ctx& c, const Level& lv0, FcnSets* s) {
  if (lv0.n == 0) return SGNN_OK;
  RC(fork_side(c, lv0.n));
  RC(coarsen_begin(c, lv0, &s->lv1));
  RC(coarsen_begin(c, s->lv1, &s->lv2));
  Level* pend[2] = {&s->lv1, &s->lv2};
  RC(counts_post(c, pend, 2));
  use_main(c);
  return SGNN_OK;
}

static int fcn(Ctx& c, const Level& lv0, FcnSets& sets, const SgnnFcnW& f, const float* x_raw, const float* x_bn, float** out,
               int64_t rows[3], int lvl) {
  const int ch = f.c;
  ALLOC(J0, float, lv0.n * 3 * ch);
  *out = J0;
  rows[0] = lv0.n; rows[1] = rows[2] = 0;
  if (lv0.n == 0) return SGNN_OK;
  Level& lv1 = sets.lv1;
  Level& lv2 = sets.lv2;
  int32_t *par01, *chi01, *par12 = nullptr, *chi12 = nullptr;
  Level* pend[2] = {&lv1, &lv2};
  // the finest residual block does not depend on the coarse sets: it is enqueued before the host asks for their counts
  ALLOC(y0_bn, float, lv0.n * ch);
  RC(res_block(c, lv0, f.blk[0], ch, x_raw, x_bn, epi_bn(J0, 3 * ch, f.bn_join), epi_bn(y0_bn, ch, f.bn_down[0])));
  RC(mark(c, "FCN residual block, finest", lvl));
  RC(counts_wait(c, pend, 2));
  use_side(c);
  RC(coarsen_finish(c, lv0, &lv1, &par01, &chi01, true, ch));
  RC(coarsen_finish(c, lv1, &lv2, &par12, &chi12, true, ch));
  RC(join_side(c));
  RC(mark(c, "FCN coarse coords + rulebooks + plans", lvl));
  rows[1] = lv1.n;
  rows[2] = lv2.n;
  ALLOC(J1, float, lv1.n * 2 * ch);
  if (lv1.n) {
    ALLOC(z1_raw, float, lv1.n * ch);
    ALLOC(z1_bn, float, lv1.n * ch);
    RC(conv(c, y0_bn, ch, ch, chi01, lv1.n, 8, 0, f.w_down[0], ch, lv1.n, nullptr, 0, epi(z1_raw, ch),
            epi_bn(z1_bn, ch, f.blk[1].bn0), lv0.n));
    ALLOC(y1_bn, float, lv1.n * ch);
    RC(res_block(c, lv1, f.blk[1], ch, z1_raw, z1_bn, epi(J1, 2 * ch), epi_bn(y1_bn, ch, f.bn_down[1])));
    ALLOC(y2, float, lv2.n * ch);
    if (lv2.n) {
      ALLOC(z2_raw, float, lv2.n * ch);
      ALLOC(z2_bn, float, lv2.n * ch);
      RC(conv(c, y1_bn, ch, ch, chi12, lv2.n, 8, 0, f.w_down[1], ch, lv2.n, nullptr, 0, epi(z2_raw, ch),
              epi_bn(z2_bn, ch, f.blk[2].bn0), lv1.n));
      RC(res_block(c, lv2, f.blk[2], ch, z2_raw, z2_bn, epi(y2, ch), kNoEpi));
    }
    SgnnEpilogue e1;
    e1.out = J1 + ch; e1.ld = 2 * ch; e1.relu = 0; e1.scale = nullptr; e1.shift = nullptr;
    RC(sgnn_unpool(y2, ch, par12, ch, lv1.n, &e1, c.stream));
  }
  RC(mark(c, "FCN coarse convolutions + inner unpool", lvl));
  SgnnEpilogue e0;
  e0.out = J0 + ch; e0.ld = 3 * ch; e0.relu = 1; e0.scale = f.bn_join.scale + ch; e0.shift = f.bn_join.shift + ch;
  RC(sgnn_unpool(J1, 2 * ch, par01, 2 * ch, lv0.n, &e0, c.stream));
  return mark(c, "FCN outer unpool", lvl);
}

// leading dimension of the joined feature rows: a multiple of 8 floats keeps every row 32-byte aligned (256-bit gathers)
static int pad8(int v) { return (v + 7) / 8 * 8; }

struct Skip { SgnnGrid g; const float* f; int c; int64_t n; };

}  // namespace

extern "C" size_t sgnn_generator_prepared_bytes(const SgnnGeneratorW* w) {
  if (!w) return 0;
  ConvW cv[64];
  const int n = enumerate_convs(w, cv);
  size_t b = 0;
  for (int i = 0; i < n; ++i)
    if (bank_eligible(cv[i])) b += bank_bytes(cv[i]);
  return b;
}

extern "C" int sgnn_generator_prepare(const SgnnGeneratorW* w, void* stream) {
  if (!w || !w->prepared) return SGNN_E_INVALID;
  if (w->prepared_bytes < sgnn_generator_prepared_bytes(w)) return SGNN_E_NOMEM;
  ConvW cv[64];
  const int n = enumerate_convs(w, cv);
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    if (!bank_eligible(cv[i])) continue;
    RC(sgnn_conv_tc32_prepare(cv[i].w, cv[i].K, cv[i].cin, cv[i].cout, cv[i].child, (char*)w->prepared + off, bank_half(cv[i]), stream));
    if (cv[i].child)
      RC(sgnn_conv_urc_prepare(cv[i].w, cv[i].cin, (char*)w->prepared + off + bank_half(cv[i]), bank_half(cv[i]), stream));
    off += bank_bytes(cv[i]);
  }
  return SGNN_OK;
}

extern "C" int sgnn_generator_forward(const SgnnGeneratorW* w, const void* coords, int coords_i64,
                                      const float* feats, int64_t n, int32_t nb, const int32_t* dims3, void* arena,
                                      size_t arena_bytes, int flags, SgnnGeneratorOut* out, void* stream) {
  if (!w || !out || !dims3 || n < 0 || nb <= 0 || (n > 0 && (!coords || !feats)) || !arena) return SGNN_E_INVALID;
  memset(out, 0, sizeof(*out));
  Ctx c;
  c.ar.base = (char*)arena; c.ar.cap = arena_bytes; c.ar.off = 0; c.ar.high = 0; c.ar.oom = false;
  c.st = (cudaStream_t)stream; c.stream = stream;
  c.main_st = c.st;
  c.side_on = false;
  c.overlap_max_rows = w->overlap_max_rows > 0 ? w->overlap_max_rows : (w->overlap_max_rows < 0 ? 0 : SGNN_DEFAULT_OVERLAP_MAX_ROWS);
  c.side_st = c.overlap_max_rows > 0 ? side_stream() : nullptr;
  c.profile = (flags & SGNN_GEN_PROFILE) != 0; c.n_ev = 0;
  c.tc32 = (flags & SGNN_GEN_TC32) != 0;
  c.compact = (flags & SGNN_GEN_DENSE_RULES) == 0;
  c.phases = (flags & SGNN_GEN_PHASES) != 0; c.n_ph = 0;
  c.w = w;
  c.tc32_min_rows = w->tc32_min_rows > 0 ? w->tc32_min_rows : SGNN_DEFAULT_TC32_MIN_ROWS;
  c.ur_min_rows = w->ur_min_rows > 0 ? w->ur_min_rows : SGNN_DEFAULT_UR_MIN_ROWS;
  int rc = SGNN_OK;
  do {
#define GEN(call)                 \
  if ((rc = (call)) != SGNN_OK) break
#define GALLOC(var, type, count)                                \
  type* var = (type*)c.ar.get((size_t)(count) * sizeof(type));  \
  if (!var) { rc = SGNN_E_NOMEM; break; }
    int dims[3] = {dims3[0], dims3[1], dims3[2]};
    // ------------------------------------------------------------ encoder (model.py:145-150)
    Level lv;
    GALLOC(status, int32_t, 1);
    if ((rc = (cudaMemsetAsync(status, 0, 4, c.st) == cudaSuccess ? SGNN_OK : SGNN_E_CUDA)) != SGNN_OK) break;
    GEN(mark(c, "start"));
    GEN(build_level(c, coords, coords_i64, n, nb, dims, status, &lv, c.compact));
    GEN(mark(c, "encoder: input grid + rulebook"));
    out->rows[0] = n;
    Skip skips[4];
    const float* x = feats;
    int ld_x = w->enc[0].cin;
    // the whole encoder pyramid (3 coarse site sets + rulebooks) up front: one host read instead of three mid-stream
    Level enc_lv[4];
    int32_t *enc_par[3], *enc_chi[3];
    enc_lv[0] = lv;
    GEN(coarsen_begin(c, enc_lv[0], &enc_lv[1]));
    GEN(coarsen_begin(c, enc_lv[1], &enc_lv[2]));
    GEN(coarsen_begin(c, enc_lv[2], &enc_lv[3]));
    int32_t* pinb = nullptr;
    GEN(pinned_slots(&pinb));
    if ((rc = (cudaMemcpyAsync(pinb + 8, status, 4, cudaMemcpyDeviceToHost, c.st) == cudaSuccess ? SGNN_OK : SGNN_E_CUDA)) !=
        SGNN_OK)
      break;
    Level* enc_pend[3] = {&enc_lv[1], &enc_lv[2], &enc_lv[3]};
    GEN(counts_post(c, enc_pend, 3));
    GEN(mark(c, "encoder: pyramid masks + scans"));
    bool enc_ok = true;
    for (int l = 0; l < 3 && enc_ok; ++l) {
      const SgnnEncLevelW& e = w->enc[l];
      const int ch = e.c;
      enc_ok = false;
      GALLOC(a_raw, float, lv.n * ch);
      GALLOC(a_bn, float, lv.n * ch);
      GEN(conv(c, x, ld_x, e.cin, lv.nbr, lv.n, 27, 0, e.w_in, ch, lv.n, nullptr, 0, epi(a_raw, ch),
               epi_bn(a_bn, ch, e.res.bn0), 0, lv.plan, lv.slots, lv.cnt));
      GALLOC(skip, float, lv.n * ch);
      GEN(res_block(c, lv, e.res, ch, a_raw, a_bn, epi_bn(skip, ch, e.bn_out), kNoEpi));
      skips[l].g = lv.g; skips[l].f = skip; skips[l].c = ch; skips[l].n = lv.n;
      if (l == 0) {
        // the three convolutions above ran on the input level while the host waited for the pyramid's counts (out-of-range
        // coordinates were masked out of the site set by sgnn_grid_build, so that work was safe; it is discarded here)
        GEN(counts_wait(c, enc_pend, 3));
        if (pinb[8]) { rc = SGNN_E_INVALID; break; }   // a coordinate outside [0, dims) x [0, nb): scn raises here too
        GEN(coarsen_finish(c, enc_lv[0], &enc_lv[1], &enc_par[0], &enc_chi[0], true, w->enc[1].c));
        GEN(coarsen_finish(c, enc_lv[1], &enc_lv[2], &enc_par[1], &enc_chi[1], true, w->enc[2].c));
        GEN(mark(c, "encoder: input-level convolutions"));
        GEN(coarsen_finish(c, enc_lv[2], &enc_lv[3], &enc_par[2], &enc_chi[2], false));
        GEN(mark(c, "encoder: pyramid coords + rulebooks + plans"));
      }
      Level cl = enc_lv[l + 1];
      int32_t* chi = enc_chi[l];
      out->rows[l + 1] = cl.n;
      GALLOC(h, float, cl.n * ch);
      if (cl.n) {
        GEN(conv(c, skip, ch, ch, chi, cl.n, 8, 0, e.w_down, ch, cl.n, nullptr, 0, epi_bn(h, ch, e.bn_down), kNoEpi, lv.n));
      }
      lv = cl;
      x = h;
      ld_x = ch;
      enc_ok = true;
    }
    if (rc || !enc_ok) { if (!rc) rc = SGNN_E_NOMEM; break; }
    GEN(mark(c, "encoder: remaining convolutions"));
    skips[3].g = lv.g; skips[3].f = x; skips[3].c = ld_x; skips[3].n = lv.n;   // ft3 (model.py:64)
    int dd[3] = {lv.dims[0], lv.dims[1], lv.dims[2]};
    // ------------------------------------------------------------ dense U-Net (model.py:152-166)
    const int cd = ld_x;
    const int64_t dvol = (int64_t)dd[0] * dd[1] * dd[2];
    GALLOC(dense, float, (int64_t)nb * cd * dvol);
    GEN(sgnn_sparse_to_dense(x, cd, lv.coords, lv.n, cd, dense, nb, dd[0], dd[1], dd[2], stream));
    const float* lay_out[6];
    int lay_c[6], lay_d[6][3];
    const float* cur = dense;
    int cur_c = cd, cur_d[3] = {dd[0], dd[1], dd[2]};
    bool dense_ok = true;
    for (int l = 0; l < 6 && dense_ok; ++l) {
      const SgnnDenseLayerW& L = w->dense[l];
      dense_ok = false;
      const float* in1 = nullptr;
      int c1 = 0;
      if (L.cat_with >= 0) { in1 = lay_out[L.cat_with]; c1 = lay_c[L.cat_with]; }
      int od[3];
      for (int i = 0; i < 3; ++i)
        od[i] = L.transposed ? (cur_d[i] - 1) * L.stride - 2 * L.pad + L.ksize
                             : (cur_d[i] + 2 * L.pad - L.ksize) / L.stride + 1;
      GALLOC(o, float, (int64_t)nb * L.cout * od[0] * od[1] * od[2]);
      if (L.transposed) {
        GEN(sgnn_dense_convT3d(cur, cur_c, in1, c1, nb, cur_d[0], cur_d[1], cur_d[2], L.w, L.cout, L.ksize, L.stride,
                               L.pad, L.bn.scale, L.bn.shift, 1, o, stream));
      } else {
        GEN(sgnn_dense_conv3d(cur, cur_c, in1, c1, nb, cur_d[0], cur_d[1], cur_d[2], L.w, L.cout, L.ksize, L.stride,
                              L.pad, L.bn.scale, L.bn.shift, 1, o, stream));
      }
      lay_out[l] = o; lay_c[l] = L.cout;
      for (int i = 0; i < 3; ++i) lay_d[l][i] = od[i];
      cur = o; cur_c = L.cout;
      for (int i = 0; i < 3; ++i) cur_d[i] = od[i];
      dense_ok = true;
    }
    if (rc || !dense_ok) { if (!rc) rc = SGNN_E_NOMEM; break; }
    if (cur_d[0] != dd[0] || cur_d[1] != dd[1] || cur_d[2] != dd[2] || cur_c != w->nf_coarse) { rc = SGNN_E_INVALID; break; }
    GALLOC(occsdf, float, (int64_t)nb * 2 * dvol);
    GEN(sgnn_dense_conv3d(cur, cur_c, nullptr, 0, nb, dd[0], dd[1], dd[2], w->w_heads, 2, 1, 1, 0, nullptr, nullptr, 0,
                          occsdf, stream));
    GEN(mark(c, "dense U-Net (8 launches)"));
    // ------------------------------------------------------------ a8: dense -> sparse (model.py:315-336)
    const int64_t ncell = (int64_t)nb * dvol;
    GALLOC(cand0, float, ncell * 2);
    int64_t m = 0;
    int32_t* locs = nullptr;
    float* fts = nullptr;
    int ld_f = 0, live_c = 0;
    {
      GALLOC(dflags, uint8_t, ncell);
      GALLOC(offs, int32_t, ncell + 1);
      const size_t sb = sgnn_scan_scratch_bytes(ncell);
      GALLOC(scr, char, sb);
      GEN(sgnn_dense_flags(occsdf, nb, dvol, cand0, dflags, offs, scr, sb, stream));
      int32_t cnt = 0;
      GEN(read_i32(c, offs + ncell, &cnt));
      m = cnt;
      live_c = w->nf_coarse + 2;
      ld_f = pad8(live_c + skips[3].c);
      GALLOC(l0, int32_t, m * 4);
      GALLOC(f0, float, m * ld_f);
      locs = l0; fts = f0;
      GEN(sgnn_dense_write(cur, occsdf, nb, w->nf_coarse, dd[0], dd[1], dd[2], dflags, offs, locs, fts, ld_f, stream));
    }
    GEN(mark(c, "dense -> sparse (flags, scan, host read, write)"));
    out->n_cand[0] = ncell; out->cand[0] = cand0;
    out->cand_locs[0] = nullptr;   // level 0 candidates are ALL cells in batch-major raster order: implicit
    // ------------------------------------------------------------ refinement levels (model.py:387-396)
    int rdims[3] = {dd[0], dd[1], dd[2]};
    bool ref_ok = true;
    bool joined = false;     // the skip columns of `fts` were written with the rows (sgnn_heads_write_join)
    for (int h = 0; h < 3 && ref_ok; ++h) {
      const SgnnRefineW& R = w->ref[h];
      ref_ok = false;
      if (m == 0) { ref_ok = true; continue; }
      const Skip& sk = skips[3 - h];
      if (sk.n && !joined) { GEN(sgnn_concat_skip(&sk.g, sk.f, sk.c, sk.c, locs, m, fts, ld_f, live_c, stream)); }
      // (rows of an empty skip set keep the zeros written with the row)
      live_c += sk.c;
      if (live_c != R.cin) { rc = SGNN_E_INVALID; break; }
      GEN(mark(c, "skip join (level 0 only)", h));
      Level rl;
      const int ch = R.c;
      GEN(build_level(c, locs, 0, m, nb, rdims, nullptr, &rl, false, ch));
      GEN(mark(c, "grid + rulebook + plan", h));
      FcnSets rsets;
      GEN(fcn_begin(c, rl, &rsets));
      GEN(mark(c, "coarse masks + scans", h));
      GALLOC(a_raw, float, m * ch);
      GALLOC(a_bn, float, m * ch);
      GEN(conv(c, fts, ld_f, R.cin, rl.nbr, m, 27, 0, R.w_in, ch, m, nullptr, 0, epi(a_raw, ch),
               epi_bn(a_bn, ch, R.fcn.blk[0].bn0), 0, rl.plan));
      float* J0 = nullptr;
      GEN(mark(c, "first convolution", h));
      GEN(fcn(c, rl, rsets, R.fcn, a_raw, a_bn, &J0, &out->rows[4 + 3 * h], h));
      // a9: 8 children per site, never materialised: n1 in child mode + n2, heads, mask, compaction
      const int64_t ncand = 8 * m;
      GALLOC(xc, float, ncand * ch);
      GEN(conv(c, J0, 3 * ch, 3 * ch, rl.nbr, m, 27, 1, R.w_up, ch, ncand, nullptr, 0, epi_bn(xc, ch, R.bn_up), kNoEpi, m, rl.plan));
      GEN(mark(c, "child-mode convolution", h));
      GALLOC(cand, float, ncand * 2);
      GALLOC(flg, uint8_t, ncand);
      GALLOC(offs, int32_t, ncand + 1);
      const size_t sb = sgnn_scan_scratch_bytes(ncand);
      GALLOC(scr, char, sb);
      GEN(sgnn_heads_flags(xc, ch, ch, R.w_occ, R.b_occ, R.w_sdf, R.b_sdf, ncand, cand, flg, offs, scr, sb, stream));
      int32_t cnt = 0;
      GEN(read_i32(c, offs + ncand, &cnt));
      GEN(mark(c, "heads + scan + host read", h));
      out->n_cand[h + 1] = ncand; out->cand[h + 1] = cand;
      if (flags & SGNN_GEN_CAND_PARENTS) {
        out->cand_locs[h + 1] = locs;          // the parents: their 8 children each are the candidates, in this order
      } else if (flags & SGNN_GEN_CAND_LOCS) {
        GALLOC(cl, int32_t, ncand * 4);
        GEN(sgnn_children_coords(locs, m, cl, stream));
        out->cand_locs[h + 1] = cl;
      }
      const int next_skip = h < 2 ? skips[2 - h].c : skips[0].c;
      const int ld_n = pad8(ch + 2 + next_skip);
      GALLOC(nl, int32_t, (int64_t)cnt * 4);
      GALLOC(nf, float, (int64_t)cnt * ld_n);
      const Skip& nsk = h < 2 ? skips[2 - h] : skips[0];
      joined = nsk.n > 0;
      if (joined) {
        GEN(sgnn_heads_write_join(xc, ch, ch, cand, locs, ncand, flg, offs, nl, nf, ld_n, &nsk.g, nsk.f, nsk.c, nsk.c, stream));
      } else {
        GEN(sgnn_heads_write(xc, ch, ch, cand, locs, ncand, flg, offs, nl, nf, ld_n, stream));
      }
      GEN(mark(c, "candidate coords + kept rows (+ skip join)", h));
      locs = nl; fts = nf; ld_f = ld_n; live_c = ch + 2; m = cnt;
      for (int i = 0; i < 3; ++i) rdims[i] *= 2;
      ref_ok = true;
    }
    if (rc || !ref_ok) { if (!rc) rc = SGNN_E_NOMEM; break; }
    // ------------------------------------------------------------ surface prediction (model.py:398-402)
    out->n_out = m;
    out->out_locs = locs;
    if (m > 0) {
      const SgnnSurfaceW& S = w->surf;
      const Skip& sk = skips[0];
      if (sk.n && !joined) { GEN(sgnn_concat_skip(&sk.g, sk.f, sk.c, sk.c, locs, m, fts, ld_f, live_c, stream)); }
      live_c += sk.c;
      if (live_c != S.cin) { rc = SGNN_E_INVALID; break; }
      Level sl;
      const int ch = S.c;
      GEN(build_level(c, locs, 0, m, nb, rdims, nullptr, &sl, false, ch));
      GEN(mark(c, "grid + rulebook + plan", 3));
      FcnSets ssets;
      GEN(fcn_begin(c, sl, &ssets));
      GEN(mark(c, "coarse masks + scans", 3));
      GALLOC(a_raw, float, m * ch);
      GALLOC(a_bn, float, m * ch);
      GEN(conv(c, fts, ld_f, S.cin, sl.nbr, m, 27, 0, S.w_in, ch, m, nullptr, 0, epi(a_raw, ch),
               epi_bn(a_bn, ch, S.fcn.blk[0].bn0), 0, sl.plan));
      float* J0 = nullptr;
      GEN(mark(c, "first convolution", 3));
      GEN(fcn(c, sl, ssets, S.fcn, a_raw, a_bn, &J0, &out->rows[13], 3));
      GALLOC(sdf, float, m);
      GEN(sgnn_linear(J0, 3 * ch, S.w_lin, S.b_lin, sdf, 1, m, 3 * ch, 1, stream));
      out->out_sdf = sdf;
      GEN(mark(c, "TSDF head", 3));
    }
  } while (0);
  if (c.profile && c.n_ev > 0 && rc == SGNN_OK) {
    SGNN_CUDA(cudaStreamSynchronize(c.st));
    double ms = 0;
    for (int i = 0; i < c.n_ev; i += 2) {
      float t = 0.f;
      SGNN_CUDA(cudaEventElapsedTime(&t, g_ev[i], g_ev[i + 1]));
      ms += t;
      g_rec[i / 2].ms = t;
    }
    g_nrec = c.n_ev / 2;
    out->conv_ms = ms;
    out->n_conv = c.n_ev / 2;
  }
  // whatever happened above, nothing stays in flight on the side stream that the caller's stream does not wait for
  if (c.side_st) {
    use_main(c);
    if (order_after(c.main_st, c.side_st) != SGNN_OK && rc == SGNN_OK) rc = SGNN_E_CUDA;
  }
  g_nph = 0;
  if (c.phases && c.n_ph > 1 && rc == SGNN_OK) {
    SGNN_CUDA(cudaStreamSynchronize(c.st));
    for (int i = 1; i < c.n_ph; ++i) {
      float t = 0.f;
      SGNN_CUDA(cudaEventElapsedTime(&t, g_ph_ev[i - 1], g_ph_ev[i]));
      g_ph[i].ms = t;
    }
    g_nph = c.n_ph;
  }
  out->arena_used = c.ar.off;
  out->arena_needed = c.ar.high > arena_bytes ? 2 * c.ar.high : c.ar.high;
  if (c.ar.oom && rc == SGNN_OK) rc = SGNN_E_NOMEM;
  if (c.ar.oom) rc = SGNN_E_NOMEM;
  return rc;
}

// i-th convolution of the calling thread's last SGNN_GEN_PROFILE pass: rec = {n_out, cin, cout, K, child_mode, tensor-core}
extern "C" int sgnn_generator_profile_entry(int32_t i, int64_t* rec6, float* ms) {
  if (i < 0 || i >= g_nrec || !rec6 || !ms) return SGNN_E_INVALID;
  const ConvRec& r = g_rec[i];
  rec6[0] = r.n_out; rec6[1] = r.cin; rec6[2] = r.cout; rec6[3] = r.K; rec6[4] = r.child; rec6[5] = r.tc;
  *ms = r.ms;
  return SGNN_OK;
}

// i-th phase of the calling thread's last SGNN_GEN_PHASES pass (i = 1 .. n-1; entry 0 is the start mark): name (<= 47 chars) and
// the CUDA-event time since the previous mark.  SGNN_E_INVALID past the end.
extern "C" int sgnn_generator_phase_entry(int32_t i, char* name48, float* ms) {
  if (i < 0 || i >= g_nph || !name48 || !ms) return SGNN_E_INVALID;
  memcpy(name48, g_ph[i].name, sizeof(g_ph[i].name));
  *ms = g_ph[i].ms;
  return SGNN_OK;
}
