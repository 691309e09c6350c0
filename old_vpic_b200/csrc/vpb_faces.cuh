// vpb_faces.cuh -- boundary planes of the field array: local boundary conditions
// (src/field_advance/standard/local.c) and face messages (remote.c), shared by the
// field kernels.  A face is (axis X, side s); Y=(X+1)%3, Z=(X+2)%3.
#pragma once
#include "vpb_common.cuh"

namespace vpb {

enum MsgKind { MSG_GHOST_TANG_B = 0, MSG_GHOST_NORM_E, MSG_GHOST_DIV_B, MSG_SYNC_JF, MSG_SYNC_RHO, MSG_SYNC_TEB };
enum AdjKind { ADJ_TANG_E = 0, ADJ_NORM_B, ADJ_DIV_E, ADJ_JF, ADJ_RHOF, ADJ_RHOB };

// local.c:260-444: enforce a local boundary condition on the boundary plane itself
void faces_local_adjust(vpb_domain_t *dom, vpb_field_t *d_f, AdjKind which);

// Fill one kind of ghost plane on all six faces: faces with a local field bc get
// local.c:50-222; faces shared with a rank (possibly this one: periodic) get the
// message exchange of remote.c:61-279 (pack -> transport -> unpack).
void faces_ghost_exchange(vpb_domain_t *dom, vpb_field_t *d_f, MsgKind kind);

// remote.c:298-621: x pass, then y, then z over the shared planes.  d_err (may be
// NULL) accumulates the squared desynchronisation for MSG_SYNC_TEB.
void faces_sync_passes(vpb_domain_t *dom, vpb_field_t *d_f, MsgKind kind, double *d_err);

// number of floats of a face message (remote.c BEGIN_RECV sizes)
int faces_message_floats(const DomainDev &g, MsgKind kind, int face);

}  // namespace vpb
