// vpb_faces.cu -- boundary planes of the field array (see vpb_faces.cuh).
//
// local.c and remote.c spell out six faces x several loop shapes by macro; here
// every plane operation is a small descriptor (box, component, rule) and ONE
// launch executes all descriptors of a phase (blockIdx.y = descriptor), because
// these planes are tiny (<= a few hundred KB) and launch latency, not bandwidth,
// is what they cost.  Message layouts are exactly the reference's, so a message
// packed here could be unpacked by the reference and vice versa.
#include "vpb_comm.cuh"
#include "vpb_faces.cuh"

namespace vpb {

enum { cEX = 0, cEY, cEZ, cDIVE, cCBX, cCBY, cCBZ, cDIVB, cTCAX, cTCAY, cTCAZ, cRHOB, cJFX, cJFY, cJFZ, cRHOF };

struct Box { int lo[3], hi[3]; };

// plane X==p, Y in 1..nY+eY, Z in 1..nZ+eZ (local.c:30-44)
static Box plane(const DomainDev &g, int X, int p, int eY, int eZ) {
  const int n[3] = {g.nx, g.ny, g.nz};
  const int Y = (X + 1) % 3, Z = (X + 2) % 3;
  Box b;
  b.lo[X] = b.hi[X] = p;
  b.lo[Y] = 1; b.hi[Y] = n[Y] + eY;
  b.lo[Z] = 1; b.hi[Z] = n[Z] + eZ;
  return b;
}
static long box_cells(const Box &b) {
  return (long)(b.hi[0] - b.lo[0] + 1) * (b.hi[1] - b.lo[1] + 1) * (b.hi[2] - b.lo[2] + 1);
}
static int face_bc(const DomainDev &g, int X, int s) {
  int ijk[3] = {0, 0, 0};
  ijk[X] = s;
  return g.bc[VPB_BOUNDARY(ijk[0], ijk[1], ijk[2])];
}
static bool is_local(const DomainDev &g, int bc) { return bc < 0 || bc > g.nproc; }    // local.c:70
static bool is_remote(const DomainDev &g, int bc) { return bc >= 0 && bc < g.nproc; }  // grid_comm.c:17

// ---------------------------------------------------------------------------
// local boundary-condition descriptors
// ---------------------------------------------------------------------------
enum { L_COPY = 0, L_NEG, L_EXTRAP, L_ZERO, L_DOUBLE, L_HIGDON };

struct LocalOp {
  Box box;
  int kind, comp;
  int ecompT, ecompX, flip, s;     // Higdon only
  long in, to_face, tstride;       // voxel offsets
  float decay, drive, cdtX, cdtT;  // Higdon only
};
constexpr int kMaxOps = 24;
struct LocalOps { int n; LocalOp op[kMaxOps]; };

#define FV(f, v, c) FCOMP(f, g, v, c)

__device__ __forceinline__ size_t box_voxel(const Box &b, long cell, const DomainDev &g) {
  const int wx = b.hi[0] - b.lo[0] + 1, wy = b.hi[1] - b.lo[1] + 1;
  const int x = b.lo[0] + (int)(cell % wx);
  const long r = cell / wx;
  const int y = b.lo[1] + (int)(r % wy), z = b.lo[2] + (int)(r / wy);
  return (size_t)x + (size_t)g.sx * ((size_t)y + (size_t)g.sy * z);
}

__global__ void __launch_bounds__(128) local_ops_kernel(vpb_field_t *__restrict__ f, const LocalOps ops, const DomainDev g) {
  const LocalOp &o = ops.op[blockIdx.y];
  const long n = (long)(o.box.hi[0] - o.box.lo[0] + 1) * (o.box.hi[1] - o.box.lo[1] + 1) * (o.box.hi[2] - o.box.lo[2] + 1);
  for (long c = (long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long)gridDim.x * blockDim.x) {
    const long v = (long)box_voxel(o.box, c, g);
    float r;
    switch (o.kind) {
    case L_COPY: r = FV(f, v + o.in, o.comp); break;
    case L_NEG: r = -FV(f, v + o.in, o.comp); break;
    case L_EXTRAP: r = 2 * FV(f, v + o.in, o.comp) - FV(f, v + 2 * o.in, o.comp); break;
    case L_ZERO: r = 0; break;
    case L_DOUBLE: r = FV(f, v, o.comp) * 2; break;
    default: {  // first-order Higdon absorber, local.c:84-111
      const long vh = v + o.in, vf = v + o.to_face;
      float t1 = o.cdtX * (FV(f, vf + o.in, o.ecompT) - FV(f, vf, o.ecompT));
      t1 = o.s < 0 ? t1 : -t1;
      float t2 = FV(f, vh + o.tstride, o.ecompX);
      t2 = o.cdtT * (t2 - FV(f, vh, o.ecompX));
      const float base = o.decay * FV(f, v, o.comp) + o.drive * FV(f, vh, o.comp);
      r = o.flip ? (base + t1 - t2) : (base - t1 + t2);
    }
    }
    FV(f, v, o.comp) = r;
  }
}

static void run_ops(vpb_field_t *d_f, LocalOps &ops, const DomainDev &g) {
  if (ops.n == 0) return;
  long maxc = 1;
  for (int i = 0; i < ops.n; i++) { long c = box_cells(ops.op[i].box); if (c > maxc) maxc = c; }
  int gx = (int)((maxc + 127) / 128);
  if (gx > 1024) gx = 1024;
  local_ops_kernel<<<dim3(gx, ops.n), 128, 0, ctx().stream>>>(d_f, ops, g);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

static void add_op(LocalOps &ops, const Box &b, int kind, int comp, long in = 0) {
  if (box_cells(b) <= 0) return;
  if (ops.n >= kMaxOps) VPB_ERROR("too many boundary descriptors");
  LocalOp &o = ops.op[ops.n++];
  memset(&o, 0, sizeof(o));
  o.box = b; o.kind = kind; o.comp = comp; o.in = in;
}

void faces_local_adjust(vpb_domain_t *dom, vpb_field_t *d_f, AdjKind which) {
  const DomainDev &g = dom->d;
  const int n[3] = {g.nx, g.ny, g.nz};
  LocalOps ops;
  ops.n = 0;
  // Boundary planes of different axes share edges, and "double it" is not idempotent: the two faces
  // of one axis (disjoint planes) go into one launch, the three axes into consecutive launches.
  for (int X = 0; X < 3; X++) {
    for (int s = -1; s <= 1; s += 2) {
      const int Y = (X + 1) % 3, Z = (X + 2) % 3;
      const int bc = face_bc(g, X, s);
      if (!is_local(g, bc)) continue;
      const int fp = s < 0 ? 1 : n[X] + 1;
      const bool pec = bc == vpb_pec_fields, sym = bc == vpb_symmetric_fields, pmc = bc == vpb_pmc_fields,
                 absb = bc == vpb_absorb_fields;
      if (!(pec || sym || pmc || absb)) VPB_ERROR("Bad boundary condition encountered.");
      switch (which) {
      case ADJ_TANG_E:   // local.c:232-262
        if (pec) {
          add_op(ops, plane(g, X, fp, 0, 1), L_ZERO, cEX + Y); add_op(ops, plane(g, X, fp, 0, 1), L_ZERO, cTCAX + Y);
          add_op(ops, plane(g, X, fp, 1, 0), L_ZERO, cEX + Z); add_op(ops, plane(g, X, fp, 1, 0), L_ZERO, cTCAX + Z);
        }
        break;
      case ADJ_NORM_B:   // :264-290
        if (sym) add_op(ops, plane(g, X, fp, 0, 0), L_ZERO, cCBX + X);
        break;
      case ADJ_DIV_E:    // :292-318
        if (pec || absb) add_op(ops, plane(g, X, fp, 1, 1), L_ZERO, cDIVE);
        break;
      case ADJ_JF:       // :325-353
        add_op(ops, plane(g, X, fp, 0, 1), pec ? L_ZERO : L_DOUBLE, cJFX + Y);
        add_op(ops, plane(g, X, fp, 1, 0), pec ? L_ZERO : L_DOUBLE, cJFX + Z);
        break;
      case ADJ_RHOF:     // :361-387
        add_op(ops, plane(g, X, fp, 1, 1), pec ? L_ZERO : L_DOUBLE, cRHOF);
        break;
      case ADJ_RHOB:     // :394-420
        if (pec) add_op(ops, plane(g, X, fp, 1, 1), L_ZERO, cRHOB);
        break;
      }
    }
    if (which == ADJ_JF || which == ADJ_RHOF) { run_ops(d_f, ops, g); ops.n = 0; }
  }
  run_ops(d_f, ops, g);
}

static void add_local_ghost_ops(LocalOps &ops, const DomainDev &g, MsgKind kind, int face) {
  const int n[3] = {g.nx, g.ny, g.nz};
  const long st[3] = {1, g.sx, g.sxy};
  const int X = face % 3, s = face < 3 ? -1 : 1, Y = (X + 1) % 3, Z = (X + 2) % 3;
  const int bc = face_bc(g, X, s);
  const int ghost = s < 0 ? 0 : n[X] + 1, fp = s < 0 ? 1 : n[X] + 1;
  const long in = -s * st[X];
  const bool pec = bc == vpb_pec_fields, symlike = bc == vpb_symmetric_fields || bc == vpb_pmc_fields,
             absb = bc == vpb_absorb_fields;
  if (!(pec || symlike || absb)) VPB_ERROR("Bad boundary condition encountered.");
  if (kind == MSG_GHOST_TANG_B) {   // local.c:50-122
    const Box bY = plane(g, X, ghost, 1, 0), bZ = plane(g, X, ghost, 0, 1);
    if (!absb) {
      add_op(ops, bY, pec ? L_COPY : L_NEG, cCBX + Y, in);
      add_op(ops, bZ, pec ? L_COPY : L_NEG, cCBX + Z, in);
    } else {
      const float cdt[3] = {g.cvac * g.dt * g.rdx, g.cvac * g.dt * g.rdy, g.cvac * g.dt * g.rdz};
      const float higend = (float)((g.nx > 1 || g.ny > 1 || g.nz > 1) ? 1.03527618 : 1.);
      float drive = cdt[X] * higend;
      const float decay = (1 - drive) / (1 + drive);
      drive = 2 * drive / (1 + drive);
      for (int k = 0; k < 2; k++) {
        const int before = ops.n;
        add_op(ops, k ? bZ : bY, L_HIGDON, cCBX + (k ? Z : Y), in);
        if (ops.n == before) continue;
        LocalOp &o = ops.op[ops.n - 1];
        o.ecompT = cEX + (k ? Y : Z);
        o.ecompX = cEX + X;
        o.flip = k;
        o.s = s;
        o.to_face = (long)(fp - ghost) * st[X];
        o.tstride = st[k ? Y : Z];
        o.decay = decay; o.drive = drive; o.cdtX = cdt[X]; o.cdtT = cdt[k ? Y : Z];
      }
    }
  } else if (kind == MSG_GHOST_NORM_E) {   // local.c:128-176
    const Box b = plane(g, X, ghost, 1, 1);
    const int k = pec ? L_COPY : (symlike ? L_NEG : L_EXTRAP);
    add_op(ops, b, k, cEX + X, in);
    add_op(ops, b, k, cTCAX + X, in);
  } else if (kind == MSG_GHOST_DIV_B) {    // local.c:178-214
    add_op(ops, plane(g, X, ghost, 0, 0), pec ? L_COPY : (symlike ? L_NEG : L_ZERO), cDIVB, in);
  }
}

// ---------------------------------------------------------------------------
// face messages
// ---------------------------------------------------------------------------
struct Seg { Box box; int c0, c1, off, count; };
struct Plan { int nseg, header, nfloats; Seg seg[3]; };

static void add_seg(Plan &p, const Box &b, int c0, int c1) {
  Seg &s = p.seg[p.nseg++];
  s.box = b; s.c0 = c0; s.c1 = c1; s.off = p.nfloats;
  s.count = (int)(box_cells(b) * (c1 >= 0 ? 2 : 1));
  if (s.count < 0) s.count = 0;
  p.nfloats += s.count;
}

// pack: the plane the message is read from; !pack: the plane it is applied to
static Plan make_plan(const DomainDev &g, MsgKind kind, int face, bool pack) {
  const int n[3] = {g.nx, g.ny, g.nz};
  const int X = face % 3, s = face < 3 ? -1 : 1, Y = (X + 1) % 3, Z = (X + 2) % 3;
  const bool ghostkind = kind <= MSG_GHOST_DIV_B;
  int p;
  if (ghostkind) p = pack ? (s < 0 ? 1 : n[X]) : (s < 0 ? 0 : n[X] + 1);
  else p = s < 0 ? 1 : n[X] + 1;
  Plan pl;
  pl.nseg = 0;
  pl.header = kind == MSG_SYNC_TEB ? 0 : 1;
  pl.nfloats = pl.header;
  switch (kind) {
  case MSG_GHOST_TANG_B: add_seg(pl, plane(g, X, p, 1, 0), cCBX + Y, -1); add_seg(pl, plane(g, X, p, 0, 1), cCBX + Z, -1); break;
  case MSG_GHOST_NORM_E: add_seg(pl, plane(g, X, p, 1, 1), cEX + X, -1); break;
  case MSG_GHOST_DIV_B: add_seg(pl, plane(g, X, p, 0, 0), cDIVB, -1); break;
  case MSG_SYNC_JF: add_seg(pl, plane(g, X, p, 0, 1), cJFX + Y, -1); add_seg(pl, plane(g, X, p, 1, 0), cJFX + Z, -1); break;
  case MSG_SYNC_RHO: add_seg(pl, plane(g, X, p, 1, 1), cRHOF, cRHOB); break;
  case MSG_SYNC_TEB:
    add_seg(pl, plane(g, X, p, 0, 0), cCBX + X, -1);
    add_seg(pl, plane(g, X, p, 0, 1), cEX + Y, cTCAX + Y);
    add_seg(pl, plane(g, X, p, 1, 0), cEX + Z, cTCAX + Z);
    break;
  }
  return pl;
}

int faces_message_floats(const DomainDev &g, MsgKind kind, int face) { return make_plan(g, kind, face, true).nfloats; }

struct Job { Plan plan; float *buf; float dX; long in; int kind; };
struct Jobs { int n; Job job[6]; double *err; };

struct Elem { size_t v; int comp; int seg; int which; };
__device__ __forceinline__ Elem decode(const Plan &p, int e, const DomainDev &g) {
  int si = 0;
  if (p.nseg > 1 && e >= p.seg[1].off) si = 1;
  if (p.nseg > 2 && e >= p.seg[2].off) si = 2;
  const Seg &s = p.seg[si];
  const int r = e - s.off;
  Elem el;
  el.seg = si;
  if (s.c1 >= 0) { el.which = r & 1; el.comp = el.which ? s.c1 : s.c0; el.v = box_voxel(s.box, r >> 1, g); }
  else { el.which = 0; el.comp = s.c0; el.v = box_voxel(s.box, r, g); }
  return el;
}

__global__ void __launch_bounds__(128) pack_kernel(const vpb_field_t *__restrict__ f, const Jobs J, const DomainDev g) {
  const Job &j = J.job[blockIdx.y];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < j.plan.nfloats; e += gridDim.x * blockDim.x) {
    if (e < j.plan.header) { j.buf[e] = j.dX; continue; }   // leading cell size (remote.c:82)
    const Elem el = decode(j.plan, e, g);
    j.buf[e] = FCOMP(const_cast<vpb_field_t *>(f), g, el.v, el.comp);
  }
}

__global__ void __launch_bounds__(128) unpack_kernel(vpb_field_t *__restrict__ f, const Jobs J, const DomainDev g) {
  const Job &j = J.job[blockIdx.y];
  const float dX = j.dX;
  float w_l = 0, w_r = 0, w_hl = 0, w_hr = 0;
  if (j.kind <= MSG_GHOST_DIV_B) {         // remote.c:107-109
    const float lw0 = j.buf[0];
    w_r = (float)((2. * dX) / (lw0 + dX));
    w_l = (lw0 - dX) / (lw0 + dX);
  } else if (j.kind == MSG_SYNC_JF || j.kind == MSG_SYNC_RHO) {   // remote.c:449-454, :570-575
    float hrw = j.buf[0], hlw = hrw + dX;
    hrw /= hlw;
    hlw = dX / hlw;
    w_hl = hlw; w_hr = hrw;
    w_l = hlw + hlw; w_r = hrw + hrw;
  }
  double err = 0;
  for (int e = j.plan.header + blockIdx.x * blockDim.x + threadIdx.x; e < j.plan.nfloats; e += gridDim.x * blockDim.x) {
    const Elem el = decode(j.plan, e, g);
    const float m = j.buf[e];
    float *dst = &FV(f, el.v, el.comp);
    switch (j.kind) {
    case MSG_GHOST_TANG_B: case MSG_GHOST_NORM_E: case MSG_GHOST_DIV_B:
      *dst = w_r * m + w_l * FV(f, (long)el.v + j.in, el.comp);
      break;
    case MSG_SYNC_JF:
      *dst = w_l * (*dst) + w_r * m;
      break;
    case MSG_SYNC_RHO:
      *dst = el.which ? (w_hl * (*dst) + w_hr * m) : (w_l * (*dst) + w_r * m);
      break;
    default: {   // MSG_SYNC_TEB, remote.c:341-372: average in double, tca does not count as error
      const double w1 = m, w2 = *dst;
      *dst = (float)(0.5 * (w1 + w2));
      if (el.which == 0) err += (w1 - w2) * (w1 - w2);
    }
    }
  }
  if (j.kind == MSG_SYNC_TEB && J.err) {
    __shared__ double ws[4];
    err = warp_sum(err);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = err;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(J.err, ws[0] + ws[1] + ws[2] + ws[3]);
  }
}

static void ensure_face_buffers(vpb_domain_t *dom) {
  const DomainDev &g = dom->d;
  for (int face = 0; face < 6; face++) {
    size_t need = 0;
    for (int k = 0; k <= MSG_SYNC_TEB; k++) {
      size_t b = (size_t)faces_message_floats(g, (MsgKind)k, face) * sizeof(float);
      if (b > need) need = b;
    }
    need = (need + 255) & ~(size_t)255;
    if (dom->face_cap[face] >= need) continue;
    if (dom->face_send[face]) { VPB_CUDA(cudaStreamSynchronize(ctx().stream)); cudaFree(dom->face_send[face]); cudaFree(dom->face_recv[face]); }
    VPB_CUDA(cudaMalloc(&dom->face_send[face], need));
    VPB_CUDA(cudaMalloc(&dom->face_recv[face], need));
    dom->face_cap[face] = need;
  }
}

static float cell_size(const DomainDev &g, int X) { return X == 0 ? g.dx : (X == 1 ? g.dy : g.dz); }

// pack the listed faces, move the messages, unpack them.  faces[] lists faces whose bc is a rank.
static void exchange_faces(vpb_domain_t *dom, vpb_field_t *d_f, MsgKind kind, const int *faces, int nf, double *d_err,
                           LocalOps *between) {
  const DomainDev &g = dom->d;
  const long st[3] = {1, g.sx, g.sxy};
  cudaStream_t stream = ctx().stream;
  if (nf > 0) {
    ensure_face_buffers(dom);
    Jobs P;
    P.n = nf; P.err = nullptr;
    int maxf = 1;
    for (int i = 0; i < nf; i++) {
      Job &j = P.job[i];
      j.plan = make_plan(g, kind, faces[i], true);
      j.buf = dom->face_send[faces[i]];
      j.dX = cell_size(g, faces[i] % 3);
      j.in = 0; j.kind = kind;
      if (j.plan.nfloats > maxf) maxf = j.plan.nfloats;
    }
    int gx = (maxf + 127) / 128;
    if (gx > 512) gx = 512;
    pack_kernel<<<dim3(gx, nf), 128, 0, stream>>>(d_f, P, g);
    count_launch();
  }
  if (between && between->n) run_ops(d_f, *between, g);
  if (nf == 0) return;
  // transport: sends by face 0..5, receives by face 3,4,5,0,1,2 (vpb_comm.cuh)
  Xfer x[12];
  int nx = 0;
  bool has[6] = {false, false, false, false, false, false};
  for (int i = 0; i < nf; i++) has[faces[i]] = true;
  for (int face = 0; face < 6; face++) {
    if (!has[face]) continue;
    const int peer = face_bc(g, face % 3, face < 3 ? -1 : 1);
    if (peer == g.rank) continue;
    Xfer &t = x[nx++];
    t.send = dom->face_send[face]; t.send_bytes = (size_t)faces_message_floats(g, kind, face) * 4; t.send_peer = peer;
    t.recv = nullptr; t.recv_bytes = 0; t.recv_peer = -1;
  }
  static const int rorder[6] = {3, 4, 5, 0, 1, 2};
  for (int k = 0; k < 6; k++) {
    const int face = rorder[k];
    if (!has[face]) continue;
    const int peer = face_bc(g, face % 3, face < 3 ? -1 : 1);
    if (peer == g.rank) continue;
    Xfer &t = x[nx++];
    t.send = nullptr; t.send_bytes = 0; t.send_peer = -1;
    // the neighbour's opposite face has the same transverse extent, hence the same size
    t.recv = dom->face_recv[face]; t.recv_bytes = (size_t)faces_message_floats(g, kind, face) * 4; t.recv_peer = peer;
  }
  if (nx) comm_exchange(x, nx);
  Jobs U;
  U.n = nf; U.err = d_err;
  int maxf = 1;
  for (int i = 0; i < nf; i++) {
    const int face = faces[i], X = face % 3, s = face < 3 ? -1 : 1;
    const int peer = face_bc(g, X, s);
    Job &j = U.job[i];
    j.plan = make_plan(g, kind, face, false);
    // what arrives through my face F is what the neighbour packed from its face (F+3)%6;
    // when the neighbour is this rank, that is my own send buffer
    j.buf = (peer == g.rank) ? dom->face_send[(face + 3) % 6] : dom->face_recv[face];
    j.dX = cell_size(g, X);
    j.in = -s * st[X];
    j.kind = kind;
    if (j.plan.nfloats > maxf) maxf = j.plan.nfloats;
  }
  int gx = (maxf + 127) / 128;
  if (gx > 512) gx = 512;
  unpack_kernel<<<dim3(gx, nf), 128, 0, stream>>>(d_f, U, g);
  count_launch();
  VPB_CUDA(cudaGetLastError());
}

void faces_ghost_exchange(vpb_domain_t *dom, vpb_field_t *d_f, MsgKind kind) {
  const DomainDev &g = dom->d;
  int faces[6], nf = 0;
  LocalOps ops;
  ops.n = 0;
  for (int face = 0; face < 6; face++) {
    const int bc = face_bc(g, face % 3, face < 3 ? -1 : 1);
    if (is_remote(g, bc)) {
      if (bc == g.rank && !is_remote(g, face_bc(g, face % 3, face < 3 ? 1 : -1)))
        VPB_ERROR("face %d is joined to this rank but the opposite face is not", face);
      faces[nf++] = face;
    } else if (is_local(g, bc)) add_local_ghost_ops(ops, g, kind, face);
  }
  exchange_faces(dom, d_f, kind, faces, nf, nullptr, &ops);
}

void faces_sync_passes(vpb_domain_t *dom, vpb_field_t *d_f, MsgKind kind, double *d_err) {
  const DomainDev &g = dom->d;
  for (int X = 0; X < 3; X++) {   // x pass, then y, then z (remote.c:281-296)
    int faces[2], nf = 0;
    // unpack order of the reference: the message from the +X side first
    if (is_remote(g, face_bc(g, X, 1))) faces[nf++] = X + 3;
    if (is_remote(g, face_bc(g, X, -1))) faces[nf++] = X;
    if (nf) exchange_faces(dom, d_f, kind, faces, nf, d_err, nullptr);
  }
}

}  // namespace vpb
