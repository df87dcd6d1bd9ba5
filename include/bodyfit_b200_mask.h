/*
 * bodyfit_b200_mask.h -- silhouette term of the SMPLify loop (use_mask=True).  Replaces smplify/loss.py:85-130
 * multview_mask_loss as called at smplify/smplify.py:196-199 (weight 5 at :210); the contours come from
 * smplify/loss.py:73-83 (cv2.findContours on the host, once per fit, as in the reference).
 * Conventions as in bodyfit_b200.h (extern "C", device pointers, caller's stream, 0 / negative code, no allocation).
 */
#ifndef BODYFIT_B200_MASK_H
#define BODYFIT_B200_MASK_H
#include <stdint.h>
#include "bodyfit_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct BfMask {
    const float*   masks;     /* [B, Nm, H, W] 0 / 1 (smplify.py:139: mask > 128) */
    const float*   cams;      /* [Nm, 12] K [R|t] of the mask views (world -> pixel) */
    const float*   contour;   /* [total, 2] (x, y) contour pixels of every (frame, view), concatenated in (frame, view) order */
    const int32_t* cptr;      /* [B*Nm + 1] offsets into contour */
    const int32_t* cown;      /* [total] frame * Nm + view of every contour pixel */
    float*         uv;        /* [B, Nm, Nq, 2] scratch: projected sampled vertices */
    int32_t*       near_q;    /* [total] scratch: closest projected in-image vertex of every contour pixel (-1: none) */
    float*         cdist;     /* [total] scratch: its distance */
    float*         cw;        /* [total] scratch: 1, or epsilon if that vertex's pixel is outside the mask */
    float*         dPw;       /* [B, Nq, 3] scratch: gradient wrt the world-space sampled vertices */
    float*         part;      /* [B, Nq] scratch: bilinear (1 - mask) samples summed over views */
    float*         mask_loss; /* [B] out (optional): the unweighted term */
    int32_t        Nm, H, W, Nq, stride, total;   /* stride = 4: every 4th vertex (loss.py:100); Nq = ceil(V / stride) */
    float          imsize, epsilon;               /* epsilon = 10 (loss.py:85) */
} BfMask;

/* all-vertex buffers of `f` (model-space f->verts of the current forward): f->loss[b] += weight * term_b, the gradient is
 * ADDED to f->dverts (model space) and f->grad[:, 0:4] (transl, scale) */
int bf_mask_loss(const BfModel* m, const BfFrames* f, const BfMask* k, float weight, void* stream);

#ifdef __cplusplus
}
#endif
#endif
