// vvb200_general.cuh -- the any-topology thermostat (included by vvb200_device.cu).
//
// The fused two-pass kernels need every thermostat molecule and every Drude pair inside one 512-slot tile.  Systems
// that cannot be tiled that way (a thermostatted polymer longer than a tile, Drude partners far apart in index order,
// molecules interleaved in particle order) run the thermostat through three gather kernels over the reference's own
// index arrays instead -- same arithmetic per particle / pair / molecule as the tiled kernels, no tile-locality
// assumption, everything still on the device (NH chains included):
//
//   general_com_kernel     one warp per thermostat molecule              [calcCOMVelocities, drudeNoseHoover.cu:5]
//   general_ke_kernel      normal particles, pairs, molecules; block-reduced group energies / bias moments, last
//                          block advances the NH chains                  [computeNormalizedKineticEnergies :55,
//                          sumNormalizedKineticEnergies :121, calcPeriodicVelocityBias + sumV cosineAccelerate.cu:16,34]
//   general_scale_kernel   in-place scaling of normal particles and pairs, bias remove / restore
//                          [normalizeVelocities :37, scaleVelocity :157, cosineAccelerate.cu:63,76]
//
// The kick runs as pass A with its molecule / pair phases switched off (KParams::kickOnly) on plain 512-slot tiles;
// drifts and the position write use the element-wise delta / finish kernels.
#pragma once

struct GParams {
    int N, nMolNH, nNormal, nPairs;
    const int32_t *moleculesNH, *normalNH, *particleMolId, *sortedByMol, *particlesInMolecules;
    const int2 *pairsNH;
    void *velm;
    const void *posq;
    void *comV, *comCbar;
    double *partials;
    NhcDevice *nhc;
    unsigned int *counter;
    double dt, invBoxZ;
    int useCOM, cosine, fuseNHC;
};

template <int MODE>
__global__ void __launch_bounds__(256) general_com_kernel(const GParams p) {
    typedef Prec<MODE> P;
    typedef typename P::real real;
    typedef typename P::mixed mixed;
    typedef typename P::real4 real4;
    typedef typename P::mixed4 mixed4;
    const mixed4 *velm = reinterpret_cast<const mixed4 *>(p.velm);
    const real4 *posq = reinterpret_cast<const real4 *>(p.posq);
    mixed4 *comV = reinterpret_cast<mixed4 *>(p.comV);
    mixed *comCbar = reinterpret_cast<mixed *>(p.comCbar);
    const int lane = threadIdx.x & 31;
    const int warpsPerGrid = (blockDim.x >> 5) * gridDim.x;
    const real invBoxZ = (real) p.invBoxZ;
    for (int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < p.nMolNH; j += warpsPerGrid) {
        const int mol = p.moleculesNH[j];
        const int cnt = p.particlesInMolecules[2 * mol], start = p.particlesInMolecules[2 * mol + 1];
        mixed sx = 0, sy = 0, sz = 0, sc = 0, comMass = 0;
        for (int k = lane; k < cnt; k += 32) {
            const int i = p.sortedByMol[start + k];
            const mixed4 v = velm[i];
            if (v.w != 0) {
                const mixed mass = vv_recip(v.w);
                sx += v.x * mass; sy += v.y * mass; sz += v.z * mass;
                if (p.cosine) sc += cosPhase((double) posq[i].z, (double) invBoxZ) * mass;
                comMass += mass;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, off);
            sy += __shfl_xor_sync(0xffffffffu, sy, off);
            sz += __shfl_xor_sync(0xffffffffu, sz, off);
            sc += __shfl_xor_sync(0xffffffffu, sc, off);
            comMass += __shfl_xor_sync(0xffffffffu, comMass, off);
        }
        if (lane == 0) {
            mixed4 V;
            V.w = vv_recip(comMass);
            V.x = sx * V.w; V.y = sy * V.w; V.z = sz * V.w;
            comV[mol] = V;
            if (p.cosine) comCbar[mol] = sc * V.w;
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) general_ke_kernel(const GParams p) {
    typedef Prec<MODE> P;
    typedef typename P::real real;
    typedef typename P::mixed mixed;
    typedef typename P::real4 real4;
    typedef typename P::mixed4 mixed4;
    __shared__ double red[8][VVB200_NRED];
    __shared__ unsigned int ticket;
    const mixed4 *velm = reinterpret_cast<const mixed4 *>(p.velm);
    const real4 *posq = reinterpret_cast<const real4 *>(p.posq);
    const mixed4 *comV = reinterpret_cast<const mixed4 *>(p.comV);
    const mixed *comCbar = reinterpret_cast<const mixed *>(p.comCbar);
    const real invBoxZ = (real) p.invBoxZ;
    const bool cosine = p.cosine;
    const int tid = threadIdx.x, stride = blockDim.x * gridDim.x, first = blockIdx.x * blockDim.x + tid;

    mixed acc[VVB200_NRED];
#pragma unroll
    for (int k = 0; k < VVB200_NRED; k++) acc[k] = 0;

    // normal particles (drudeNoseHoover.cu:76-83)
    for (int k = first; k < p.nNormal; k += stride) {
        const int i = p.normalNH[k];
        const mixed4 v = velm[i];
        if (v.w == 0) continue;
        const mixed mass = vv_recip(v.w);
        mixed Vx = 0, Vy = 0, Vz = 0, cb = 0;
        if (p.useCOM) {
            const int mol = p.particleMolId[i];
            const mixed4 V = comV[mol];
            Vx = V.x; Vy = V.y; Vz = V.z;
            if (cosine) cb = comCbar[mol];
        }
        const mixed ux = v.x - Vx, uy = v.y - Vy, uz = v.z - Vz;
        acc[0] += (ux * ux + uy * uy + uz * uz) * mass;
        if (cosine) {
            const mixed d = cosPhase((double) posq[i].z, (double) invBoxZ) - cb;
            acc[4] += ux * d * mass;
            acc[7] += d * d * mass;
        }
    }
    // Drude pairs (drudeNoseHoover.cu:99-114)
    for (int k = first; k < p.nPairs; k += stride) {
        const int2 pr = p.pairsNH[k];
        const mixed4 v1 = velm[pr.x], v2 = velm[pr.y];
        mixed Vx = 0, Vy = 0, Vz = 0, cb = 0;
        if (p.useCOM) {
            const int mol = p.particleMolId[pr.x];
            const mixed4 V = comV[mol];
            Vx = V.x; Vy = V.y; Vz = V.z;
            if (cosine) cb = comCbar[mol];
        }
        const mixed mass1 = vv_recip(v1.w), mass2 = vv_recip(v2.w);
        const mixed u1x = v1.x - Vx, u1y = v1.y - Vy, u1z = v1.z - Vz;
        const mixed u2x = v2.x - Vx, u2y = v2.y - Vy, u2z = v2.z - Vz;
        const mixed totalMass = mass1 + mass2;
        const mixed invTotalMass = vv_recip(totalMass);
        const mixed m1f = invTotalMass * mass1, m2f = invTotalMass * mass2;
        const mixed redMass = mass1 * m2f;
        const mixed cmx = u1x * m1f + u2x * m2f, cmy = u1y * m1f + u2y * m2f, cmz = u1z * m1f + u2z * m2f;
        const mixed rx = u1x - u2x, ry = u1y - u2y, rz = u1z - u2z;
        acc[0] += (cmx * cmx + cmy * cmy + cmz * cmz) * totalMass;
        acc[2] += (rx * rx + ry * ry + rz * rz) * redMass;
        if (cosine) {
            const mixed d1 = cosPhase((double) posq[pr.x].z, (double) invBoxZ) - cb;
            const mixed d2 = cosPhase((double) posq[pr.y].z, (double) invBoxZ) - cb;
            const mixed cmd = d1 * m1f + d2 * m2f, rd = d1 - d2;
            acc[4] += cmx * cmd * totalMass;
            acc[7] += cmd * cmd * totalMass;
            acc[6] += rx * rd * redMass;
            acc[9] += rd * rd * redMass;
        }
    }
    // molecular group (drudeNoseHoover.cu:91-97)
    if (p.useCOM)
        for (int j = first; j < p.nMolNH; j += stride) {
            const int mol = p.moleculesNH[j];
            const mixed4 V = comV[mol];
            if (V.w != 0) {
                const mixed M = vv_recip(V.w);
                acc[1] += (V.x * V.x + V.y * V.y + V.z * V.z) * M;
                if (cosine) {
                    const mixed cb = comCbar[mol];
                    acc[5] += M * V.x * cb;
                    acc[8] += M * cb * cb;
                }
            }
        }
    // velocity bias over ALL massive atoms (cosineAccelerate.cu:16-32)
    if (cosine)
        for (int i = first; i < p.N; i += stride) {
            const mixed4 v = velm[i];
            if (v.w != 0)
                acc[3] += vv_recip(v.w) * v.x * 2 * cosPhase((double) posq[i].z, (double) invBoxZ);
        }

    // block reduction, last block sums the partials in a fixed order and advances the chains (as in pass A)
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < VVB200_NRED; k++) {
        double v = (double) acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (tid < VVB200_NRED) {
        double v = 0;
        for (int w = 0; w < (int) (blockDim.x >> 5); w++) v += red[w][tid];
        p.partials[(size_t) blockIdx.x * VVB200_NRED + tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0)
        ticket = atomicAdd(p.counter, 1u);
    __syncthreads();
    if (ticket != gridDim.x - 1)
        return;
    __threadfence();
    for (int k = 0; k < VVB200_NRED; k++) {
        double v = 0;
        for (int b = tid; b < (int) gridDim.x; b += blockDim.x)
            v += __ldcg(p.partials + (size_t) b * VVB200_NRED + k);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (tid < VVB200_NRED) {
        double v = 0;
        for (int w = 0; w < (int) (blockDim.x >> 5); w++) v += red[w][tid];
        p.nhc->red[tid] = v;
    }
    if (tid == 0)
        *p.counter = 0;
    __syncthreads();
    if (p.fuseNHC && tid < 3) {
        if (cosine) nhcFinish<true>(p.nhc, p.dt, tid);
        else nhcFinish<false>(p.nhc, p.dt, tid);
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) general_scale_kernel(const GParams p) {
    typedef Prec<MODE> P;
    typedef typename P::real real;
    typedef typename P::mixed mixed;
    typedef typename P::real4 real4;
    typedef typename P::mixed4 mixed4;
    mixed4 *velm = reinterpret_cast<mixed4 *>(p.velm);
    const real4 *posq = reinterpret_cast<const real4 *>(p.posq);
    const mixed4 *comV = reinterpret_cast<const mixed4 *>(p.comV);
    const mixed *comCbar = reinterpret_cast<const mixed *>(p.comCbar);
    const real invBoxZ = (real) p.invBoxZ;
    const bool cosine = p.cosine;
    const mixed sA = (mixed) p.nhc->vscale[0], sC = (mixed) p.nhc->vscale[1], sD = (mixed) p.nhc->vscale[2];
    const mixed Vb = cosine ? (mixed) p.nhc->vBias : (mixed) 0;
    const int stride = blockDim.x * gridDim.x, first = blockIdx.x * blockDim.x + threadIdx.x;

    for (int k = first; k < p.nNormal; k += stride) {
        const int i = p.normalNH[k];
        mixed4 v = velm[i];
        double c = 0;
        if (cosine) { c = cosPhase((double) posq[i].z, (double) invBoxZ); v.x -= Vb * c; }
        mixed Vn[3] = {0, 0, 0};
        if (p.useCOM) {
            const int mol = p.particleMolId[i];
            const mixed4 V = comV[mol];
            Vn[0] = V.x; Vn[1] = V.y; Vn[2] = V.z;
            if (cosine) Vn[0] = V.x - Vb * comCbar[mol];
            v.x -= Vn[0]; v.y -= Vn[1]; v.z -= Vn[2];
        }
        if (v.w != 0) {
            v.x = sA * v.x + sC * Vn[0];
            v.y = sA * v.y + sC * Vn[1];
            v.z = sA * v.z + sC * Vn[2];
        }
        if (cosine) v.x += Vb * c;
        velm[i] = v;
    }
    for (int k = first; k < p.nPairs; k += stride) {
        const int2 pr = p.pairsNH[k];
        mixed4 a = velm[pr.x], b = velm[pr.y];
        double c1 = 0, c2 = 0;
        if (cosine) {
            c1 = cosPhase((double) posq[pr.x].z, (double) invBoxZ);
            c2 = cosPhase((double) posq[pr.y].z, (double) invBoxZ);
            a.x -= Vb * c1;
            b.x -= Vb * c2;
        }
        mixed Vn[3] = {0, 0, 0};
        if (p.useCOM) {
            const int mol = p.particleMolId[pr.x];
            const mixed4 V = comV[mol];
            Vn[0] = V.x; Vn[1] = V.y; Vn[2] = V.z;
            if (cosine) Vn[0] = V.x - Vb * comCbar[mol];
        }
        mixed v1[3] = {a.x - Vn[0], a.y - Vn[1], a.z - Vn[2]}, v2[3] = {b.x - Vn[0], b.y - Vn[1], b.z - Vn[2]};
        const mixed invW = vv_recip(a.w + b.w);
        mixed o1[3], o2[3];
        scalePair<mixed>(v1, v2, b.w * invW, a.w * invW, Vn, sA, sC, sD, o1, o2);
        a.x = o1[0]; a.y = o1[1]; a.z = o1[2];
        b.x = o2[0]; b.y = o2[1]; b.z = o2[2];
        if (cosine) { a.x += Vb * c1; b.x += Vb * c2; }
        velm[pr.x] = a;
        velm[pr.y] = b;
    }
}
