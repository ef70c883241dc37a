// grid.h -- HBM data layout shared by the host code and the kernels.
//
// Cell fields (density, mask): the reference layout, c = x + y*W + (z-zlo)*W*H.  The derived stencil-code
// array has its own padded pitch, k = x + y*PC + (z-zlo)*PC*H, so 2- and 16-byte accesses are aligned.
// Staggered fields (u, v, w): "node" layout with a padded row pitch so that every row starts on a
// 32-byte sector and TMA / float4 access is legal: n = x + y*P + (z-zlo)*P*(H+1), P = roundup(W+1, 8).
// Node (x,y,z) holds the three LOW faces of cell (x,y,z): u[x,y,z] between cells x-1|x, v between
// y-1|y, w between z-1|z (MAC convention of the reference).  z is the slowest index, so a
// z-slab [zlo, zlo+nz) of any field is one contiguous range (multi-GPU halos need no packing).
#pragma once
#include <cstdint>

struct GridP {
    int W, H, D;      // global cell counts
    int P;            // node row pitch in floats
    int SY;           // node rows per plane = H + 1
    long long nplane; // node plane stride  = P * SY
    long long cplane; // cell plane stride  = W * H
    int zlo;          // global z of the first stored plane (0 on a single GPU)
    int nzc;          // stored cell planes  [zlo, zlo + nzc)
    int nzn;          // stored node planes  [zlo, zlo + nzn), nzn = nzc + 1
    int PC;           // row pitch of the stencil-code array in bytes = roundup(W, 16) (pad bytes are 0)
    long long kplane; // plane stride of the stencil-code array = PC * H
    int mzlo;         // first stored MASK plane: a slab keeps one more mask plane than cell planes on each interior
    int nzm;          // side so that the stencil codes of its outermost stored cells are right; [mzlo, mzlo + nzm)
};

// stencil code byte per cell, derived from the mask after every fill
enum : unsigned {
    CODE_SX0 = 1u, CODE_SX1 = 2u, CODE_SY0 = 4u, CODE_SY1 = 8u, CODE_SZ0 = 16u, CODE_SZ1 = 32u,
    CODE_ACTIVE = 64u, // interior fluid cell with at least one fluid neighbour: updated by the pressure sweeps
    CODE_SELF = 128u   // the cell itself is fluid
};

// Second encoding of the same information for the fused pressure passes ("pcode", one byte per cell, same pitch):
// bits 0-5 as above, bit 6 = ACTIVE, bit 7 = COMPLEX for an ACTIVE cell (fewer than six fluid neighbours: the update
// needs the per-face masks and its own neighbour count) and = SELF for a cell that is not ACTIVE.  So
//   (p & 0xC0) == 0x40   ACTIVE with six fluid neighbours  -> acc = 6, every face updated (the common case),
//   (p & 0xC0) == 0xC0   ACTIVE next to a solid / the domain shell -> general update,
//   (p & 0xC0) != 0      the cell is fluid (what the forcing term needs, cu:321),
// and a warp decides "every cell of this sweep is ACTIVE and simple" with one LOP3 + one vote.
enum : unsigned { PCODE_ACTIVE = 64u, PCODE_COMPLEX = 128u };

#define SMK_MAX_OBJ 16
struct ObjP {
    int nsrc, nobs;
    int obstacle_union; // 0: the last obstacle decides (reference), 1: solid inside any obstacle (extension)
    float src[SMK_MAX_OBJ][4]; // x, y, z, r           (cu:721-727)
    float obs[SMK_MAX_OBJ][4]; // x, y, z, r  (the obstacle velocity is never read by a kernel, cu:299-301)
};
