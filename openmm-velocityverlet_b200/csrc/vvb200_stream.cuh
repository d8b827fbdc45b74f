// vvb200_stream.cuh -- the two streaming kernels of the fused integrator step (included by vvb200_device.cu).
//
// Both are persistent, warp-specialised kernels built around a TMA (cp.async.bulk) -> shared memory ring:
//
//   warp 8 (producer)   for every molecule-aligned tile of this block: waits for a free stage, arms the stage's
//                       "full" mbarrier with the byte count and issues one bulk copy per input array
//                       (velm, force x/y/z, posq, posqCorrection, the packed slot words, per-molecule tables);
//   warps 0-7 (consumers, 256 threads, 2 particles each) wait on "full", do the arithmetic from shared memory,
//                       write results straight from registers with 128/256-bit coalesced stores, and release the
//                       stage through the "empty" mbarrier.
//
// The ring keeps `stages` x ~30-40 KB of loads in flight per block regardless of what the consumers are doing
// (molecule COM phase, block barriers), which is what an HBM-bound kernel with ~100 B/particle and non-trivial
// per-tile synchronisation needs to approach the copy roofline (Little: ~35 KB in flight per SM).
//
// A tile [t0,t1) starts at an arbitrary particle index; bulk copies need 16-byte alignment, so every array is
// copied for the slot range [t0 & ~3, roundup4(t1)) and consumers index with the offset t0 - (t0 & ~3).
#pragma once

#ifndef CTHREADS
#define CTHREADS 256                          // consumer threads
#endif
#define BTHREADS (CTHREADS + 32)              // + one producer warp
#ifndef MINBLOCKS_A
#define MINBLOCKS_A 3                         // pass A: 3 co-resident blocks (<= 72 registers) hide its block barriers
#endif
// pass B: ONE block per SM with 16 consumer warps (one particle per thread) and a 4-deep ring.  Measured against two
// blocks of 8 consumer warps with 2 stages each (same warps, same shared memory): 341.6 vs 362 us for 16.4M particles
// -- a tile is finished in half the time, so the stage goes back to the producer sooner and three tiles instead of
// one are in flight behind the one being computed.
// The scale-only variant (64 B/particle, little arithmetic) is the exception: two independent 8-warp blocks overlap
// one block's wait with the other's stores better than one wide block does (156 vs 178 us), so the consumer count is a
// template parameter of the kernel and chosen per variant (passBConsumers).
#ifndef MAXSTAGES_B
#define MAXSTAGES_B 4
#endif
__host__ __device__ constexpr int passBConsumers(int variant) { return variant == VAR_SCALE_ONLY ? 256 : 512; }
__host__ __device__ constexpr int passBMinBlocks(int variant) { return variant == VAR_SCALE_ONLY ? 2 : 1; }
#define ITEMS ((VVB200_TILE_CAP + CTHREADS - 1) / CTHREADS)   // particles per consumer thread (the last round may be partial)
#define PADT (VVB200_TILE_CAP + 8)            // stage slots: tile + alignment slack
#define MAXMOL VVB200_TILE_MAX_MOLS
static_assert(VVB200_TILE_CAP <= 1024, "11-bit tile-local molecule ids");

// -DVVB200_TRACE (diagnostic builds only, tests/diag_trace.py): %globaltimer stamps of thread 0 of every block
#ifdef VVB200_TRACE
__device__ unsigned long long g_vvb200Trace[1024 * 8];
__device__ __forceinline__ void traceMark(int i) {
    if (threadIdx.x == 0 && blockIdx.x < 1024) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g_vvb200Trace[blockIdx.x * 8 + i] = t;
    }
}
// streaming kernels: one row per block and pass (rows 0..1023 pass A, 1024..2047 pass B), stamped by the calling thread
__device__ unsigned long long g_vvb200TraceS[2048 * 8];
__device__ __forceinline__ void traceMarkS(int pass, int i) {
    if (blockIdx.x < 1024) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g_vvb200TraceS[(pass * 1024 + blockIdx.x) * 8 + i] = t;
    }
}
#else
#define traceMark(i) ((void) 0)
#define traceMarkS(pass, i) ((void) 0)
#endif

// ---- mbarrier / bulk-copy PTX ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fenceBarrierInit() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbarArriveExpectTx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smemAddr(bar)), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: sleep, do not spin
}
// global -> shared bulk copy (TMA engine, SASS UBLKCP), completion counted in bytes on `bar`
__device__ __forceinline__ void bulkLoad(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dst)),
                 "l"(src), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}
// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may be
// scheduled while its predecessor in the stream is still draining; it must not touch the predecessor's data before
// gridDepWait() (which returns once that grid has completed and its writes are visible).  gridDepLaunch() tells the
// runtime that this block no longer needs to hold the successor back.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void gridDepWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gridDepLaunch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void consumerBarrier() { asm volatile("bar.sync 1, %0;" ::"n"(CTHREADS) : "memory"); }

// ---- early hand-over from pass A to pass B ------------------------------------------------------------------------
// griddepcontrol.wait in pass B would wait for ALL of pass A: the last block's sum over the per-block partials, the whole
// chain update, its write-back and the grid's completion -- 4-5 us in which the other 147 SMs have nothing to do, a
// tenth of a 1M-particle step.  Pass B needs two things only: every tile's kicked velocities (done once the last
// arrival ticket is taken) and the three scale factors (done after nhcCrit).  So
//   * every block of pass A tells the runtime right after ITS OWN griddepcontrol.wait that dependents may be scheduled:
//     pass B's blocks take over an SM as soon as pass A's blocks there have left (they cannot co-reside: registers),
//     and since a dependent grid starts only after ALL blocks of pass A have said so, every block of pass A is resident
//     or done by then -- a waiting pass B can never keep a block of pass A from running;
//   * block 0 of pass A zeroes the hand-over word and the factor records before that (after its wait, i.e. after the
//     previous pass B is completely done with them); the last block sets the word to 1 (velocities visible) and, right
//     after nhcCrit, writes each factor as a self-validating 16-byte record (device.cu, FactorRec);
//   * ONE thread per block of pass B polls the word (it has a 128-byte line to itself; thousands of pollers on the line
//     of the arrival ticket slowed pass A's own tail down): the producer waits for >= 1 before its first bulk copy,
//     and fills the ring; four lanes of the first consumer warp poll the factor records, put the values into shared
//     memory and open a barrier the other warps sleep on (bounded waits: on expiry the factors are poisoned with NaN,
//     the device never hangs);
//   * pass B's blocks start at different times now (the SM that hosts pass A's last block comes last), so under the
//     hand-over they draw their tiles from a counter instead of striding -- from the END of the system backwards: what
//     pass A wrote last is what L2 still holds.  (Pass A keeps its static stride: its per-block partial sums must not
//     depend on timing.)
// p.flagSync is set by the entry points that launch both passes themselves; everything else keeps griddepcontrol.wait.
__device__ __forceinline__ void passAHandOverStart(const KParams &p, const int tid) {
    if (tid == 0) {
        if (p.flagSync && blockIdx.x == 0) {
            *p.tileTicket = 0u;
            for (int k = 0; k < 4; k++) factorPublish(p.factorRec + k, 0.0, 0ull);
            st_release_gpu(p.syncFlag, 0u);
            __threadfence();
        }
        gridDepLaunch();
    }
}

// tile descriptor (two int4 per tile, built by vvb200_plan_upload)
//   d0 = (t0, t1, m0, nMol)   d1 = (molFirst or -1, 0, 0, 0)
// molFirst >= 0: the tile's thermostat molecules are the consecutive ids molFirst .. molFirst+nMol-1

template <int MODE, bool EXTRA, bool FORCE = true> struct StageA {      // FORCE = false: the reduce-only variant (KICK_NONE)
    typename Prec<MODE>::mixed4 velm[PADT];
    long long f[3][FORCE ? PADT : 2];
    uint32_t meta[PADT];
    int32_t molInfo[MAXMOL + 8];
    int32_t desc[8];
    typename Prec<MODE>::real4 posq[EXTRA ? PADT : 1];
};

// what phase 1 publishes for the molecule and pair phases: kicked velocities and masses as a structure of
// arrays (conflict-free shared-memory access), double-buffered so that the next tile's phase 1 may start while
// slow warps still read this tile's.  The TMA stage itself is released right after phase 1.
template <int MODE, bool EXTRA> struct PublishedA {
    typedef typename Prec<MODE>::mixed mixed;
    mixed vx[VVB200_TILE_CAP], vy[VVB200_TILE_CAP], vz[VVB200_TILE_CAP], m[VVB200_TILE_CAP];
    double cph[EXTRA ? VVB200_TILE_CAP : 2];
    int32_t molInfo[MAXMOL];
};

template <int MODE, bool EXTRA> struct ScratchA {
    typedef typename Prec<MODE>::mixed mixed;
    PublishedA<MODE, EXTRA> pub[2];
    double red[CTHREADS / 32][VVB200_NRED];
    unsigned long long peerSeq;
    unsigned int ticket;
};

__host__ __device__ constexpr size_t roundUp128(size_t x) { return (x + 127) / 128 * 128; }

template <int MODE, bool EXTRA, bool FORCE> constexpr size_t smemBytesA(int stages) {
    return roundUp128(sizeof(StageA<MODE, EXTRA, FORCE>)) * stages + roundUp128(sizeof(ScratchA<MODE, EXTRA>)) + 16 * stages + 128;
}

// co-resident blocks per SM pass A is compiled for.  The reduce-only variant (no forces staged, no kick) is bound by
// instruction latency, not bytes, and would fit a fourth block -- but only at 56 registers per thread, and the spills
// cost more than the extra warps bring (252 vs 226 us for 16.4M particles), so it stays at three like the others.
#ifndef MINBLOCKS_A_REDUCE
#define MINBLOCKS_A_REDUCE MINBLOCKS_A
#endif
__host__ __device__ constexpr int passABlocks(int kick) { return kick == KICK_NONE ? MINBLOCKS_A_REDUCE : MINBLOCKS_A; }

// Lanes that cooperate on one molecule's centre of mass.  The butterfly that adds their partial sums up costs
// log2(COM_LANES) x 8 shuffles per group, and shuffles share the shared-memory (LSU) pipe that bounds these kernels:
// 4 lanes instead of 8 took the reduce-only pass from 176 to 149 us for 16.4M particles and left pass A where it was
// (242 us).  The same value everywhere: the group partition decides which thread adds a molecule's M|V|^2, and the
// kernels' sums are meant to be bit-identical.
#ifndef COM_LANES
#define COM_LANES 4
#endif

// ------------------------------------------------------------------------------------------------
// pass A: extra forces + kick + molecular COM + group kinetic energies (+ bias moments) + NH chains
// ------------------------------------------------------------------------------------------------
// The per-tile work is written once (passAPhase1 / passAPhase23 / blockReduce / lastBlockFinish) and used by the
// streaming kernel below and by the single-launch resident kernel (vvb200_resident.cuh).  RESIDENT = the kicked
// velocities, molecular velocities and cosine means stay in the shared-memory stage for pass B of the same launch.
template <int MODE> struct ACtx {
    typedef typename Prec<MODE>::real real;
    typedef typename Prec<MODE>::mixed mixed;
    mixed stepSize, fscale;
    real efscale, accel, invBoxZ;
    bool cosine, useCOM;
};

template <int MODE, int KICK> __device__ __forceinline__ ACtx<MODE> makeACtx(const KParams &p, bool extra) {
    typedef typename Prec<MODE>::real real;
    typedef typename Prec<MODE>::mixed mixed;
    ACtx<MODE> c;
    c.stepSize = (mixed) p.dt;
    // middle.cu:11-12 / CudaVVKernels.cpp:306
    c.fscale = KICK == KICK_VV ? (mixed) (0.5 * p.dt / (double) 0x100000000) : c.stepSize / (mixed) 0x100000000;
    c.efscale = (real) p.efscale;
    c.accel = (real) p.accel;
    c.invBoxZ = (real) p.invBoxZ;
    c.cosine = extra && p.cosine;
    c.useCOM = p.useCOM;
    return c;
}

// addExtraForceDrudeLangevin (drudeLangevin.cu:2-59) for ONE particle, evaluated inside the kick instead of by a
// launch of its own: same expressions as langevin_force_kernel (which the velocity-Verlet scheme still uses, because
// its first half re-applies the force of the half before).  `v` is this particle's pre-kick velocity; a pair member
// reads its partner's from the stage (pairs are never cut by a tile boundary).
template <int MODE, class Stage>
__device__ __forceinline__ void langevinForceInline(const KParams &p, const Stage &st, const int sl, const int slot,
                                                    const uint32_t mw, const typename Prec<MODE>::mixed4 v,
                                                    typename Prec<MODE>::real &ex, typename Prec<MODE>::real &ey,
                                                    typename Prec<MODE>::real &ez) {
    typedef typename Prec<MODE>::real real;
    typedef typename Prec<MODE>::mixed mixed;
    typedef typename Prec<MODE>::mixed4 mixed4;
    const mixed dragFactor = (mixed) p.ldDrag, randFactor = (mixed) p.ldRand;
    const uint32_t role = (mw >> VVB200_META_ROLE_SHIFT) & VVB200_META_ROLE_MASK;
    if (slot < p.nNormalLD) {
        ex = ey = ez = 0;
        if (v.w != 0) {
            const mixed mass = vv_recip(v.w);
            const mixed sqrtMass = vv_sqrt<MODE, mixed>(mass);
            const float4 r = __ldg(p.random + p.randomIndex + slot);
            ex += (-dragFactor * mass * v.x + randFactor * sqrtMass * r.x);
            ey += (-dragFactor * mass * v.y + randFactor * sqrtMass * r.y);
            ez += (-dragFactor * mass * v.z + randFactor * sqrtMass * r.z);
        }
        return;
    }
    const mixed dragFactorDrude = (mixed) p.ldDragDrude, randFactorDrude = (mixed) p.ldRandDrude;
    const bool selfIsDrude = role == VVB200_ROLE_DRUDE;
    const int psl = sl + (int) (mw >> VVB200_META_PARTNER_SHIFT) - VVB200_META_PARTNER_BIAS;
    const mixed4 vp = st.velm[psl];
    const mixed4 v1 = selfIsDrude ? v : vp, v2 = selfIsDrude ? vp : v;      // (Drude, parent) like pairsLD
    const int pairSlot = selfIsDrude ? slot : slot - 1;                     // slots: nNormal + 2k (Drude), + 1 (parent)
    const mixed mass1 = vv_recip(v1.w), mass2 = vv_recip(v2.w);
    const mixed totMass = mass1 + mass2;
    const mixed sqrtTotMass = vv_sqrt<MODE, mixed>(totMass);
    const mixed redMass = vv_recip((mass1 + mass2) * v1.w * v2.w);
    const mixed sqrtRedMass = vv_sqrt<MODE, mixed>(redMass);
    const mixed invTotMass = vv_recip(totMass);
    const mixed m1f = invTotMass * mass1, m2f = invTotMass * mass2;
    const float4 r1 = __ldg(p.random + p.randomIndex + pairSlot), r2 = __ldg(p.random + p.randomIndex + pairSlot + 1);
    const mixed cm[3] = {v1.x * m1f + v2.x * m2f, v1.y * m1f + v2.y * m2f, v1.z * m1f + v2.z * m2f};
    const mixed rel[3] = {v2.x - v1.x, v2.y - v1.y, v2.z - v1.z};
    const float ra[3] = {r1.x, r1.y, r1.z}, rb[3] = {r2.x, r2.y, r2.z};
    const real f1 = (real) m1f, f2 = (real) m2f;
    real out[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const real cmForce = (real) (-dragFactor * totMass * cm[d] + randFactor * sqrtTotMass * ra[d]);
        const real relForce = (real) (-dragFactorDrude * redMass * rel[d] + randFactorDrude * sqrtRedMass * rb[d]);
        out[d] = selfIsDrude ? (real) 0 + (f1 * cmForce - relForce) : (real) 0 + (f2 * cmForce + relForce);
    }
    ex = out[0]; ey = out[1]; ez = out[2];
}

// ---- phase 1: extra forces, kick, store; publish v' and mass ----------------------------------------
template <int MODE, int KICK, bool EXTRA, bool RESIDENT, class Stage>
__device__ __forceinline__ void passAPhase1(const KParams &p, const ACtx<MODE> &cx, Stage &st, PublishedA<MODE, EXTRA> &pub,
                                            typename Prec<MODE>::mixed4 (&vel)[ITEMS],   // .w holds the MASS (0 for massless)
                                            uint32_t (&meta)[ITEMS],
                                            typename Prec<MODE>::mixed (&acc)[EXTRA ? VVB200_NRED : 3], const int tid) {
    typedef Prec<MODE> P;
    typedef typename P::real real;
    typedef typename P::mixed mixed;
    typedef typename P::real4 real4;
    typedef typename P::mixed4 mixed4;
    typedef typename P::real3 real3;
    mixed4 *velm = reinterpret_cast<mixed4 *>(p.velm);
    const real3 *ldForce = reinterpret_cast<const real3 *>(p.ldForce);
    const mixed stepSize = cx.stepSize, fscale = cx.fscale;
    const real efscale = cx.efscale, accel = cx.accel, invBoxZ = cx.invBoxZ;
    const bool cosine = cx.cosine;
    const int t0 = st.desc[0], t1 = st.desc[1], m0 = st.desc[2], nMol = cx.useCOM ? st.desc[3] : 0;
    const int sl0 = t0 - (t0 & ~3), ml0 = m0 - (m0 & ~3);
    (void) velm;
    // ---- phase 1: extra forces, kick, store; publish v' and mass --------------------------------
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        const int loc = it * CTHREADS + tid;
        const int idx = t0 + loc, sl = sl0 + loc;
        meta[it] = VVB200_META_MOL_NONE;
        vel[it].x = vel[it].y = vel[it].z = vel[it].w = 0;
        if (idx < t1) {
            meta[it] = st.meta[sl];
            mixed4 v = st.velm[sl];
            double cph = 0;
            real q = 0;
            if (EXTRA) {
                const real4 pq = st.posq[sl];
                q = pq.w;
                if (cosine) cph = cosPhase((double) pq.z, (double) invBoxZ);
            }
            if (KICK != KICK_NONE && v.w != 0) {
                const long long fx = st.f[0][sl], fy = st.f[1][sl], fz = st.f[2][sl];
                if (EXTRA) {
                    // forceExtra as the reference builds it: reset, += Langevin, += field, += cosine
                    real ex = 0, ey = 0, ez = 0;
                    if (p.extraForces) {
                        if (p.hasLD && (meta[it] & VVB200_META_LD)) {
                            if (KICK == KICK_MIDDLE && p.ldInline) {
                                langevinForceInline<MODE>(p, st, sl, p.ldSlot[idx], meta[it], v, ex, ey, ez);
                            } else {
                                const real3 f = ldForce[p.ldSlot[idx]];
                                ex = f.x; ey = f.y; ez = f.z;
                            }
                        }
                        if (p.hasField) {
                            const int cnt = (meta[it] >> VVB200_META_ELEC_SHIFT) & VVB200_META_ELEC_MASK;
                            for (int c = 0; c < cnt; c++)
                                ez += efscale * q;                         // electricField.cu:10
                        }
                        if (cosine)   // cosineAccelerate.cu:9 (float += double unless double mode)
                            ex = (real) (ex + accel * cph * vv_recip(v.w));
                    }
                    if (KICK == KICK_MIDDLE) {   // middle.cu:17-19
                        v.x += stepSize * v.w * ex + fscale * v.w * fx;
                        v.y += stepSize * v.w * ey + fscale * v.w * fy;
                        v.z += stepSize * v.w * ez + fscale * v.w * fz;
                    } else {                      // velocityVerlet.cu:19-21 (0.5 is a double literal)
                        v.x += 0.5 * stepSize * v.w * ex + fscale * v.w * fx;
                        v.y += 0.5 * stepSize * v.w * ey + fscale * v.w * fy;
                        v.z += 0.5 * stepSize * v.w * ez + fscale * v.w * fz;
                    }
                } else {
                    // forceExtra == 0: the reference's first product is an exact zero
                    v.x += fscale * v.w * fx;
                    v.y += fscale * v.w * fy;
                    v.z += fscale * v.w * fz;
                }
                // resident kernel: the kicked velocity stays on chip for pass B, but it goes back into the stage only
                // after the block barrier (passAPhase23): a Langevin pair partner may still need the pre-kick value
                if constexpr (!RESIDENT) st_stream(velm + idx, v);
            }
            const mixed mass = v.w != 0 ? rcpMass(v.w) : (mixed) 0;
            vel[it] = v;
            vel[it].w = mass;
            pub.vx[loc] = v.x; pub.vy[loc] = v.y; pub.vz[loc] = v.z; pub.m[loc] = mass;
            if (EXTRA) {
                pub.cph[loc] = cph;
                if (cosine && v.w != 0)   // cosineAccelerate.cu:26
                    acc[3] += mass * v.x * 2 * cph;
            }
            // Atom-group energy (drudeNoseHoover.cu:76-114).  The reference sums m|v - V_mol|^2 over the normal particles
            // and (m1+m2)|v_cm - V_mol|^2 over the Drude pairs; here EVERY massive thermostat particle adds m|v|^2, then
            //   - M|V_mol|^2 leaves once per molecule in phase 2 (every massive member of a thermostat molecule is in the
            //     sum), so no particle has to wait for its molecule's V, and
            //   - mu|v1 - v2|^2 leaves once per Drude pair in phase 3 and goes to the Drude group, because
            //     m1|v1|^2 + m2|v2|^2 = (m1+m2)|v_cm|^2 + mu|v1-v2|^2: the pair phase needs no pair-COM velocity at all.
            // Same numbers up to fp64 reassociation (<= 1e-15 per term; the tests hold the sums to 1e-12).
            if (!p.kickOnly && (meta[it] & VVB200_META_NH) && v.w != 0) {
                acc[0] += (v.x * v.x + v.y * v.y + v.z * v.z) * mass;
                if (cosine) {
                    acc[4] += v.x * cph * mass;
                    acc[7] += cph * cph * mass;
                }
            }
        }
    }
    if (tid < nMol) pub.molInfo[tid] = st.molInfo[ml0 + tid];
}

// ---- phases 2 and 3 (after a block barrier): molecular COM velocities, then the Drude pairs -----------
template <int MODE, bool EXTRA, bool RESIDENT, class Stage, int KICK = KICK_NONE>
__device__ __forceinline__ void passAPhase23(const KParams &p, const ACtx<MODE> &cx, Stage &st, PublishedA<MODE, EXTRA> &pub,
                                             const typename Prec<MODE>::mixed4 (&vel)[ITEMS], const uint32_t (&meta)[ITEMS],
                                             typename Prec<MODE>::mixed (&acc)[EXTRA ? VVB200_NRED : 3], const int tid,
                                             const int t0, const int t1, const int m0, const int nMol, const int molFirst) {
    typedef Prec<MODE> P;
    typedef typename P::mixed mixed;
    typedef typename P::mixed4 mixed4;
    mixed4 *comV = reinterpret_cast<mixed4 *>(p.comV);
    mixed *comCbar = reinterpret_cast<mixed *>(p.comCbar);
    const bool cosine = cx.cosine;
    (void) st;
    if constexpr (RESIDENT && KICK != KICK_NONE) {
        // the kicked velocities into the stage (each thread its own slots; read again after the grid barrier)
        const int sl0 = t0 - (t0 & ~3);
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            const int loc = it * CTHREADS + tid;
            if (t0 + loc < t1 && vel[it].w != 0) {
                st.velm[sl0 + loc].x = vel[it].x; st.velm[sl0 + loc].y = vel[it].y; st.velm[sl0 + loc].z = vel[it].z;
            }
        }
    }
    // ---- phase 2: molecular centre-of-mass velocities (drudeNoseHoover.cu:11-30): COM_LANES lanes per
    //      molecule stride over its particles, then a butterfly over the lane group (fixed order) ----
    if (nMol > 0) {
        const int grp = tid / COM_LANES, sub = tid % COM_LANES;
        for (int jb = 0; jb < nMol; jb += CTHREADS / COM_LANES) {
            if (jb + (tid & ~31) / COM_LANES >= nMol)
                break;      // warp-uniform: no lane of this warp holds a molecule, in this round or any later one
            const int j = jb + grp;
            const bool active = j < nMol;
            mixed sx = 0, sy = 0, sz = 0, sc = 0, comMass = 0;
            uint32_t info = 0;
            int mol = 0;
            if (active) {
                info = (uint32_t) pub.molInfo[j];
                mol = molFirst >= 0 ? molFirst + j : p.tileMolList[m0 + j];
                if (!MOLINFO_SCATTERED(info)) {
                    const int first = MOLINFO_FIRST(info), cnt = MOLINFO_COUNT(info);
                    for (int k = first + sub; k < first + cnt; k += COM_LANES) {
                        const mixed mass = pub.m[k];        // 0 for massless particles: no contribution
                        sx += pub.vx[k] * mass; sy += pub.vy[k] * mass; sz += pub.vz[k] * mass;
                        if (cosine) sc += pub.cph[k] * mass;
                        comMass += mass;
                    }
                } else if (sub == 0) {
                    // members interleaved with other molecules: walk the sorted list (any topology)
                    const int cnt = p.particlesInMolecules[2 * mol], start = p.particlesInMolecules[2 * mol + 1];
                    for (int k = 0; k < cnt; k++) {
                        const int loc = p.sortedByMol[start + k] - t0;
                        if (loc < 0 || loc >= t1 - t0) continue;   // massless non-thermostatted members elsewhere
                        const mixed mass = pub.m[loc];
                        sx += pub.vx[loc] * mass; sy += pub.vy[loc] * mass; sz += pub.vz[loc] * mass;
                        if (cosine) sc += pub.cph[loc] * mass;
                        comMass += mass;
                    }
                }
            }
#pragma unroll
            for (int off = COM_LANES / 2; off > 0; off >>= 1) {
                sx += __shfl_xor_sync(0xffffffffu, sx, off);
                sy += __shfl_xor_sync(0xffffffffu, sy, off);
                sz += __shfl_xor_sync(0xffffffffu, sz, off);
                comMass += __shfl_xor_sync(0xffffffffu, comMass, off);
                if (EXTRA) sc += __shfl_xor_sync(0xffffffffu, sc, off);
            }
            if (active && sub == 0 && MOLINFO_FRAGMENT(info)) {
                // part of a molecule that is longer than a tile: leave the sums for the last block (lastBlockFinish)
                double *fp = p.fragPartials + 5 * (size_t) p.tileMolFrag[m0 + j];
                fp[0] = (double) sx; fp[1] = (double) sy; fp[2] = (double) sz; fp[3] = (double) comMass; fp[4] = (double) sc;
            } else if (active && sub == 0) {
                mixed4 V;
                V.w = vv_recip(comMass);
                V.x = sx * V.w; V.y = sy * V.w; V.z = sz * V.w;
                st_stream(comV + mol, V);
                if constexpr (RESIDENT) st.comV[j] = V;
                // molecular temperature group (drudeNoseHoover.cu:91-97): |V|^2 / comVelm.w; the same amount
                // leaves the atom group (see phase 1)
                const mixed mv2 = (V.x * V.x + V.y * V.y + V.z * V.z) * comMass;
                acc[1] += mv2;
                acc[0] -= mv2;
                if (cosine) {
                    const mixed cb = sc * V.w;
                    comCbar[mol] = cb;
                    if constexpr (RESIDENT) st.cbar[j] = cb;
                    const mixed b = comMass * V.x * cb, c = comMass * cb * cb;
                    acc[5] += b; acc[4] -= b;
                    acc[8] += c; acc[7] -= c;
                }
            }
        }
    }

    // ---- phase 3: Drude pairs (drudeNoseHoover.cu:99-114), the Drude thread owning the pair: the relative-motion energy
    //      mu|v1 - v2|^2 is the Drude group's, and leaves the atom group (see phase 1) -------------------------------
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        const uint32_t mw = meta[it];
        if (!(mw & VVB200_META_NH) || ((mw >> VVB200_META_ROLE_SHIFT) & VVB200_META_ROLE_MASK) != VVB200_ROLE_DRUDE)
            continue;
        const int loc = it * CTHREADS + tid;
        const int ploc = loc + (int) (mw >> VVB200_META_PARTNER_SHIFT) - VVB200_META_PARTNER_BIAS;
        const mixed4 v = vel[it];     // .w = mass
        const mixed mass1 = v.w, mass2 = pub.m[ploc];
        const mixed redMass = mass1 * mass2 * rcpMass(mass1 + mass2);       // = 1 / ((m1+m2) w1 w2)
        const mixed rx = v.x - pub.vx[ploc], ry = v.y - pub.vy[ploc], rz = v.z - pub.vz[ploc];
        const mixed e = (rx * rx + ry * ry + rz * rz) * redMass;
        acc[2] += e;
        acc[0] -= e;
        if (cosine) {
            const mixed rd = pub.cph[loc] - pub.cph[ploc];
            const mixed b = rx * rd * redMass, c = rd * rd * redMass;
            acc[6] += b; acc[4] -= b;
            acc[9] += c; acc[7] -= c;
        }
    }
}

// NhcDevice <-> shared-memory copy by threads first .. CTHREADS - 1 (plain 8-byte words; __ldcg: the source may
// have been written by another SM earlier in this launch sequence)
static_assert(sizeof(NhcDevice) % 8 == 0, "NhcDevice copy");
__device__ __forceinline__ void nhcFetch(NhcDevice *shared, const NhcDevice *global, const int tid, const int first) {
    if (tid < first)
        return;
    for (int i = tid - first; i < (int) (sizeof(NhcDevice) / 8); i += CTHREADS - first)
        reinterpret_cast<double *>(shared)[i] = __ldcg(reinterpret_cast<const double *>(global) + i);
}
__device__ __forceinline__ void nhcStore(NhcDevice *global, const NhcDevice *shared, const int tid) {
    for (int i = tid; i < (int) (sizeof(NhcDevice) / 8); i += CTHREADS)
        reinterpret_cast<double *>(global)[i] = reinterpret_cast<const double *>(shared)[i];
}

// ---- per-block reduction of the thread accumulators (fixed order) and the arrival ticket: true for the block
//      that arrives last.  All CTHREADS consumer threads call it. -------------------------------------------------
template <int NR, class Scratch, class mixed>
__device__ __forceinline__ bool blockReduceAndTicket(const KParams &p, Scratch &sm, const mixed (&acc)[NR], const int tid) {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < NR; k++) {
        double v = (double) acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) sm.red[warp][k] = v;
    }
    consumerBarrier();
    if (tid < VVB200_NRED) {
        double v = 0;
        if (tid < NR) {
#pragma unroll
            for (int w = 0; w < CTHREADS / 32; w++) v += sm.red[w][tid];
        }
        p.partials[(size_t) blockIdx.x * VVB200_NRED + tid] = v;
    }
    __threadfence();
    consumerBarrier();
    if (tid == 0)
        sm.ticket = atomicAdd(p.counter, 1u);
    consumerBarrier();
    return sm.ticket == gridDim.x - 1;
}

// ---- the last block sums the per-block partials block-major in a fixed order and advances the NH chains ---------
// `work`: the thermostat state to advance -- p.nhc itself, or a shared-memory copy the caller prefetched and writes
// back afterwards (resident kernel: saves the chain's serial L2 round trips).
template <int MODE, int NR, class Scratch>
__device__ __forceinline__ void lastBlockFinish(const KParams &p, Scratch &sm, const bool cosine, const int tid,
                                                NhcDevice *work = nullptr, const bool splitChain = false,
                                                const int chainBase = 0, const unsigned long long tag = 1ull,
                                                const bool publish = false, const bool deferPost = false) {
    // splitChain: the chain update is cut at the scale factor (nhcCrit / nhcPost); threads chainBase .. chainBase + 2 run
    // it; publish: each factor leaves as a FactorRec with `tag` the moment it exists; deferPost: the caller runs nhcPost
    typedef typename Prec<MODE>::mixed mixed;
    typedef typename Prec<MODE>::mixed4 mixed4;
    if (!work) work = p.nhc;
    const int lane = tid & 31, warp = tid >> 5;
    // every block has stored its velocities (and whole molecules' centre-of-mass velocities) and taken its ticket:
    // pass B's producers may start loading (split molecules: after their centre of mass below).  The fence makes the
    // ticket chain an acquire for this block's loads of the partials and, with the store after it, a release for the word.
    __threadfence();
    if (p.flagSync && p.numSplit == 0 && tid == 0)
        *reinterpret_cast<volatile unsigned int *>(p.syncFlag) = 1u;      // fence + store = release: ONE fence on this path
    if (tid == 0) traceMarkS(0, 2);
    // all NR sums of a block's partials are fetched together (independent loads: one L2 round trip per block row)
    double tot[NR];
#pragma unroll
    for (int k = 0; k < NR; k++) tot[k] = 0;
    // (the first rounds unrolled: their loads are independent and leave together -- one L2 round trip, not one per round)
    {
        double part[3][NR];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int b = tid + r * CTHREADS;
#pragma unroll
            for (int k = 0; k < NR; k++)
                part[r][k] = b < (int) gridDim.x ? __ldcg(p.partials + (size_t) b * VVB200_NRED + k) : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 3; r++) {
            if (tid + r * CTHREADS < (int) gridDim.x) {
#pragma unroll
                for (int k = 0; k < NR; k++) tot[k] += part[r][k];
            }
        }
    }
    for (int b = tid + 3 * CTHREADS; b < (int) gridDim.x; b += CTHREADS) {
#pragma unroll
        for (int k = 0; k < NR; k++)
            tot[k] += __ldcg(p.partials + (size_t) b * VVB200_NRED + k);
    }
    if (tid == 0) traceMarkS(0, 3);
    // molecules cut across tiles: add their fragments up in tile order, finish the centre of mass the way phase 2 does
    // for whole molecules (drudeNoseHoover.cu:11-30, 91-97) and move M|V|^2 from the atom group to the molecular group
    // (a step cut into several launches over tile ranges -- the host pipeline -- finishes them once, in the launch that
    // covers the last tiles: by then every fragment has been written, and the M|V|^2 transfer is not added per launch)
    for (int k = tid; k < (p.kickOnly || p.tileEnd != p.numTiles ? 0 : p.numSplit); k += CTHREADS) {
        mixed sx = 0, sy = 0, sz = 0, comMass = 0, sc = 0;
        for (int f = p.splitFragOffset[k]; f < p.splitFragOffset[k + 1]; f++) {
            const double *fp = p.fragPartials + 5 * (size_t) p.splitFragList[f];
            sx += (mixed) __ldcg(fp); sy += (mixed) __ldcg(fp + 1); sz += (mixed) __ldcg(fp + 2);
            comMass += (mixed) __ldcg(fp + 3); sc += (mixed) __ldcg(fp + 4);
        }
        const int mol = p.splitMolId[k];
        mixed4 V;
        V.w = vv_recip(comMass);
        V.x = sx * V.w; V.y = sy * V.w; V.z = sz * V.w;
        reinterpret_cast<mixed4 *>(p.comV)[mol] = V;
        const mixed mv2 = (V.x * V.x + V.y * V.y + V.z * V.z) * comMass;
        tot[1] += (double) mv2;
        tot[0] -= (double) mv2;
        if (NR > 3 && cosine) {
            const mixed cb = sc * V.w;
            reinterpret_cast<mixed *>(p.comCbar)[mol] = cb;
            const mixed b = comMass * V.x * cb, c = comMass * cb * cb;
            tot[NR > 3 ? 5 : 0] += (double) b; tot[NR > 3 ? 4 : 0] -= (double) b;
            tot[NR > 3 ? 8 : 0] += (double) c; tot[NR > 3 ? 7 : 0] -= (double) c;
        }
    }
#pragma unroll
    for (int k = 0; k < NR; k++) {
        double v = tot[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) sm.red[warp][k] = v;
    }
    consumerBarrier();
    if (tid < VVB200_NRED) {
        double v = 0;
        if (tid < NR)
            for (int w = 0; w < CTHREADS / 32; w++) v += sm.red[w][tid];
        if (p.accumulateRed) v += work->red[tid];   // a step split into several launches over tile ranges (host pipeline)
        work->red[tid] = v;
        sm.red[0][tid] = v;      // thread `tid` is the only reader of column `tid`
    }
    if (tid == 0) {
        *p.counter = 0;
        if (p.flagSync && p.numSplit != 0) {      // the barrier above ordered every thread's comV stores before this
            __threadfence();
            st_release_gpu(p.syncFlag, 1u);
        }
    }
    consumerBarrier();
    if (p.peerOn) {
        // multi-GPU: this rank's sums are final -- exchange them with the peers over NVLink right here (device.cu,
        // "all-reduce of the reduction vector over NVLink peer memory"): no separate launch, and pass B still overlaps
        // its prologue with this block through PDL
        if (tid == 0) {
            sm.ticket = 0;                                   // reused as the "a wait expired" mark
            sm.peerSeq = ++*p.peer.seq;
        }
        consumerBarrier();
        const unsigned long long seq = sm.peerSeq;
        if (tid < p.peer.world && !peerPublishAndWait(p.peer, seq, sm.red[0], tid))
            sm.ticket = 1;
        consumerBarrier();
        if (tid < VVB200_NRED) {
            const double v = peerSum(p.peer, seq, tid, sm.ticket != 0);
            work->red[tid] = v;
            sm.red[0][tid] = v;
        }
        consumerBarrier();
    }
#ifdef VVB200_TRACE
    traceMark(7);
    if (tid == 0) traceMarkS(0, 7);
#endif
    const int g = tid - chainBase;
    if (p.fuseNHC && g >= 0 && g < 3) {
        if (!splitChain) {
            if (cosine) nhcFinish<true>(work, p.dt, g, sm.red[0]);
            else nhcFinish<false>(work, p.dt, g, sm.red[0]);
        } else {
            if (cosine) nhcCrit<true>(work, p.dt, g, sm.red[0]);
            else nhcCrit<false>(work, p.dt, g, sm.red[0]);
            // what pass B needs leaves right now, each factor in a self-validating record (no fence, no second word);
            // the state itself follows with nhcStore once the rest of the chain update is done
            if (publish) {
                if (p.numSplit != 0) __threadfence();      // the cut molecules' velocities (ordered by the barriers above) first
                factorPublish(p.factorRec + g, work->vscale[g], tag);
                if (g == 0) factorPublish(p.factorRec + 3, work->vBias, tag);
            }
            if (g == 0) traceMarkS(0, 5);
            if (!deferPost) {
                __syncwarp(0x7u);      // nhcPost rewrites preDt, which the other two read in nhcCrit (chainBase is a warp's first thread)
                nhcPost(work, p.dt, g);
            }
        }
    }
}

template <int MODE, int KICK, bool EXTRA>
__global__ void __launch_bounds__(BTHREADS, passABlocks(KICK)) kick_reduce_kernel(const KParams p) {
    typedef Prec<MODE> P;
    typedef typename P::real real;
    typedef typename P::mixed mixed;
    typedef typename P::real4 real4;
    typedef typename P::mixed4 mixed4;
    typedef typename P::real3 real3;
    typedef StageA<MODE, EXTRA, KICK != KICK_NONE> Stage;
    typedef ScratchA<MODE, EXTRA> Scratch;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ NhcDevice nhcS;
    const int stages = p.stagesA;
    constexpr size_t stageBytes = roundUp128(sizeof(Stage));
    Scratch &sm = *reinterpret_cast<Scratch *>(smemRaw + stageBytes * stages);
    uint64_t *full = reinterpret_cast<uint64_t *>(smemRaw + stageBytes * stages + roundUp128(sizeof(Scratch)));
    uint64_t *empty = full + stages;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < stages; s++) {
            mbarInit(full + s, 1);
            mbarInit(empty + s, CTHREADS);
        }
        fenceBarrierInit();
    }
    __syncthreads();
    if (tid == 0) traceMarkS(0, 0);
    gridDepWait();      // PDL: everything above overlapped the previous kernel's tail (pass B of the step before)
    if (tid == 0) traceMarkS(0, 1);
    passAHandOverStart(p, tid);

    const bool cosine = EXTRA && p.cosine;
    const bool useCOM = p.useCOM;

    if (tid >= CTHREADS) {
        // ===== producer warp =====
        if (tid != CTHREADS)
            return;
        int s = 0;
        uint32_t phase = 0;
        for (int tile = p.tileBegin + blockIdx.x; tile < p.tileEnd; tile += gridDim.x) {
            const int4 d0 = __ldg(p.tileDesc + 2 * tile), d1 = __ldg(p.tileDesc + 2 * tile + 1);
            mbarWait(empty + s, phase ^ 1);
            Stage &st = *reinterpret_cast<Stage *>(smemRaw + stageBytes * s);
            const int a0 = d0.x & ~3, cnt = ((d0.y + 3) & ~3) - a0;
            const int ma0 = d0.z & ~3, mcnt = useCOM && d0.w > 0 ? ((d0.z + d0.w + 3) & ~3) - ma0 : 0;
            st.desc[0] = d0.x; st.desc[1] = d0.y; st.desc[2] = d0.z; st.desc[3] = d0.w; st.desc[4] = d1.x;
            uint32_t bytes = cnt * (uint32_t) (sizeof(mixed4) + sizeof(uint32_t)) + mcnt * 4u;
            if (KICK != KICK_NONE) bytes += 3u * cnt * 8u;
            if (EXTRA) bytes += cnt * (uint32_t) sizeof(real4);
            mbarArriveExpectTx(full + s, bytes);
            bulkLoad(st.velm, reinterpret_cast<const mixed4 *>(p.velm) + a0, cnt * (uint32_t) sizeof(mixed4), full + s);
            if (KICK != KICK_NONE) {
                bulkLoad(st.f[0], p.force + a0, cnt * 8u, full + s);
                bulkLoad(st.f[1], p.force + a0 + p.paddedN, cnt * 8u, full + s);
                bulkLoad(st.f[2], p.force + a0 + 2 * (size_t) p.paddedN, cnt * 8u, full + s);
            }
            bulkLoad(st.meta, p.slotMeta + a0, cnt * 4u, full + s);
            if (mcnt) bulkLoad(st.molInfo, p.tileMolInfo + ma0, mcnt * 4u, full + s);
            if (EXTRA) bulkLoad(st.posq, reinterpret_cast<const real4 *>(p.posq) + a0, cnt * (uint32_t) sizeof(real4), full + s);
            if (++s == stages) { s = 0; phase ^= 1; }
        }
        return;
    }

    // ===== consumers =====
    // every block prefetches the thermostat state (any of them may arrive last): the chains then run on shared memory
    // instead of a string of dependent L2 round trips
    if (p.fuseNHC)
        nhcFetch(&nhcS, p.nhc, tid, 0);
    const ACtx<MODE> c = makeACtx<MODE, KICK>(p, EXTRA);
    constexpr int NR = EXTRA ? VVB200_NRED : 3;
    // per-thread accumulators in `mixed` like the reference's kineticEnergyBuffer
    mixed acc[NR];
#pragma unroll
    for (int k = 0; k < NR; k++) acc[k] = 0;

    int s = 0, buf = 0;
    uint32_t phase = 0;
    for (int tile = p.tileBegin + blockIdx.x; tile < p.tileEnd; tile += gridDim.x) {
        mbarWait(full + s, phase);
        if (tid == 0 && tile == p.tileBegin + (int) blockIdx.x) traceMarkS(0, 2);
        Stage &st = *reinterpret_cast<Stage *>(smemRaw + stageBytes * s);
        PublishedA<MODE, EXTRA> &pub = sm.pub[buf];
        const int t0 = st.desc[0], t1 = st.desc[1], m0 = st.desc[2], nMol = useCOM ? st.desc[3] : 0, molFirst = st.desc[4];
        mixed4 vel[ITEMS];
        uint32_t meta[ITEMS];
        passAPhase1<MODE, KICK, EXTRA, false>(p, c, st, pub, vel, meta, acc, tid);
        mbarArrive(empty + s);      // this thread no longer needs the stage: the producer may refill it
        if (++s == stages) { s = 0; phase ^= 1; }
        if (p.kickOnly) {           // any-topology path: the thermostat runs in the gather kernels
            buf ^= 1;
            continue;
        }
        consumerBarrier();
        passAPhase23<MODE, EXTRA, false>(p, c, st, pub, vel, meta, acc, tid, t0, t1, m0, nMol, molFirst);
        buf ^= 1;
    }

    if (tid == 0) traceMarkS(0, 3);
    if (p.kickOnly)
        return;         // nothing was reduced (vvb200_middle_kick, any-topology kick): no partials, no ticket, no last block
    if (!blockReduceAndTicket<NR>(p, sm, acc, tid)) {
        if (tid == 0) traceMarkS(0, 4);
        return;
    }
    if (tid == 0) traceMarkS(0, 4);
    if (p.fuseNHC) {
        lastBlockFinish<MODE, NR>(p, sm, cosine, tid, &nhcS, true, 0, 1ull, p.flagSync != 0);
        consumerBarrier();
        nhcStore(p.nhc, &nhcS, tid);      // the advanced state back to global memory
        if (tid == 0) traceMarkS(0, 6);
    } else {
        lastBlockFinish<MODE, NR>(p, sm, cosine, tid);
    }
}

// ------------------------------------------------------------------------------------------------
// reduce-only pass: molecular COM + group kinetic energies of the CURRENT velocities (+ NH chains)
// ------------------------------------------------------------------------------------------------
// What the thermostat needs when the kick is not part of the same call: the constraint-bearing flow (OpenMM's velocity
// constraints run between the kick and the thermostat, CudaVVKernels.cpp:151), the velocity-Verlet scheme's leading half
// step, vvb200_thermostat, vvb200_measure_temperatures.  36 B/particle (velm + slot word), nothing to write but comV.
// kick_reduce_kernel<KICK_NONE> ran it at 0.39 of the copy roofline.  This kernel does the same arithmetic in the same
// per-thread order (its sums are bit-identical to pass A's while every tile has a block of its own) with fewer
// instructions and fewer shared-memory wavefronts per particle:
//   - nothing is re-published: the molecule phase reads (vx, vy, vz, w) straight from the TMA stage and recomputes the
//     mass (rcpMass: 5 instructions);
//   - the pair phase needs only mu|v1 - v2|^2 (see passAPhase1): the Drude's lane reads its partner from the stage;
//   - with nothing written to shared memory there is NO block barrier: every warp goes from the stage's `full` barrier to
//     its `empty` arrival on its own, the warps of a block slip against each other by what the ring allows;
//   - warps whose lanes hold none of the tile's molecules leave the molecule phase before its shuffles, and the molecule
//     lanes rotate from tile to tile (below).
// What its time follows is the number of warp instructions issued (ncu source page, profiles/ncu_r02_reduce_summary.txt:
// ~2,600 per tile, issue slots 63 % busy with 27 warps per SM, fp64 pipe 48 %, DRAM 54 %): rewrites that relieved the
// shared-memory pipe (bank-conflict-free half-swapped slot loads: data-pipe utilisation 77 -> 48 %) or the dependent
// chains (straight-line code, two particles per round) but issued more instructions were all slower -- DESIGN.md 8.6c.
// Cosine runs (which also need cos(kz) per particle, published once) keep the general kernel.
template <int MODE> struct StageRed {
    typename Prec<MODE>::mixed4 velm[PADT];
    uint32_t meta[PADT];
    int32_t molInfo[MAXMOL + 8];
    int32_t desc[8];
};
struct ScratchRed {
    double red[CTHREADS / 32][VVB200_NRED];
    unsigned long long peerSeq;
    unsigned int ticket;
};
template <int MODE> constexpr size_t smemBytesRed(int stages) {
    return roundUp128(sizeof(StageRed<MODE>)) * stages + roundUp128(sizeof(ScratchRed)) + 16 * stages + 128;
}
#ifndef ROT_RED
#define ROT_RED (CTHREADS / 2)      // rotation of the molecule lanes from one tile of a block to the next (0: none)
#endif
// 3 blocks x 3 stages at 72 registers (no spills in the tile loop): with the rotated molecule lanes the deeper ring -- the
// warps slip further against each other -- is worth more than a fourth block at 56 registers: 137 -> 128 us at 16.4M
#ifndef MINBLOCKS_RED
#define MINBLOCKS_RED 3
#endif

template <int MODE>
__global__ void __launch_bounds__(BTHREADS, MINBLOCKS_RED) reduce_kernel(const KParams p) {
    typedef Prec<MODE> P;
    typedef typename P::mixed mixed;
    typedef typename P::mixed4 mixed4;
    typedef StageRed<MODE> Stage;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ NhcDevice nhcS;
    const int stages = p.stagesA;
    constexpr size_t stageBytes = roundUp128(sizeof(Stage));
    ScratchRed &sm = *reinterpret_cast<ScratchRed *>(smemRaw + stageBytes * stages);
    uint64_t *full = reinterpret_cast<uint64_t *>(smemRaw + stageBytes * stages + roundUp128(sizeof(ScratchRed)));
    uint64_t *empty = full + stages;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < stages; s++) {
            mbarInit(full + s, 1);
            mbarInit(empty + s, CTHREADS);
        }
        fenceBarrierInit();
    }
    __syncthreads();
    gridDepWait();
    passAHandOverStart(p, tid);
    const bool useCOM = p.useCOM;

    if (tid >= CTHREADS) {
        // ===== producer warp =====
        if (tid != CTHREADS)
            return;
        int s = 0;
        uint32_t phase = 0;
        for (int tile = p.tileBegin + blockIdx.x; tile < p.tileEnd; tile += gridDim.x) {
            const int4 d0 = __ldg(p.tileDesc + 2 * tile), d1 = __ldg(p.tileDesc + 2 * tile + 1);
            mbarWait(empty + s, phase ^ 1);
            Stage &st = *reinterpret_cast<Stage *>(smemRaw + stageBytes * s);
            const int a0 = d0.x & ~3, cnt = ((d0.y + 3) & ~3) - a0;
            const int ma0 = d0.z & ~3, mcnt = useCOM && d0.w > 0 ? ((d0.z + d0.w + 3) & ~3) - ma0 : 0;
            st.desc[0] = d0.x; st.desc[1] = d0.y; st.desc[2] = d0.z; st.desc[3] = d0.w; st.desc[4] = d1.x;
            mbarArriveExpectTx(full + s, cnt * (uint32_t) (sizeof(mixed4) + sizeof(uint32_t)) + mcnt * 4u);
            bulkLoad(st.velm, reinterpret_cast<const mixed4 *>(p.velm) + a0, cnt * (uint32_t) sizeof(mixed4), full + s);
            bulkLoad(st.meta, p.slotMeta + a0, cnt * 4u, full + s);
            if (mcnt) bulkLoad(st.molInfo, p.tileMolInfo + ma0, mcnt * 4u, full + s);
            if (++s == stages) { s = 0; phase ^= 1; }
        }
        return;
    }

    // ===== consumers =====
    if (p.fuseNHC)
        nhcFetch(&nhcS, p.nhc, tid, 0);
    mixed acc[3] = {0, 0, 0};
    mixed4 *comV = reinterpret_cast<mixed4 *>(p.comV);
    int s = 0, rot = 0;
    uint32_t phase = 0;
    for (int tile = p.tileBegin + blockIdx.x; tile < p.tileEnd; tile += gridDim.x) {
        mbarWait(full + s, phase);
        Stage &st = *reinterpret_cast<Stage *>(smemRaw + stageBytes * s);
        const int t0 = st.desc[0], t1 = st.desc[1], m0 = st.desc[2], nMol = useCOM ? st.desc[3] : 0, molFirst = st.desc[4];
        const int sl0 = t0 - (t0 & ~3), ml0 = m0 - (m0 & ~3);
        const mixed4 *sv = st.velm + sl0;      // (vx, vy, vz, 1/m) per tile-local slot
        // ---- per particle: m|v|^2 of every massive thermostat particle (see passAPhase1) and, on the Drude's lane, the
        //      pair's relative-motion energy mu|v1 - v2|^2 (drudeNoseHoover.cu:99-114) ----
        mixed pairRel[ITEMS];
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            const int loc = it * CTHREADS + tid;
            pairRel[it] = 0;
            if (t0 + loc < t1) {
                const mixed4 v = sv[loc];
                const uint32_t mw = st.meta[sl0 + loc];
                const mixed mass = v.w != 0 ? rcpMass(v.w) : (mixed) 0;
                if ((mw & VVB200_META_NH) && v.w != 0)
                    acc[0] += (v.x * v.x + v.y * v.y + v.z * v.z) * mass;
                if ((mw & VVB200_META_NH) && ((mw >> VVB200_META_ROLE_SHIFT) & VVB200_META_ROLE_MASK) == VVB200_ROLE_DRUDE) {
                    const mixed4 q = sv[loc + (int) (mw >> VVB200_META_PARTNER_SHIFT) - VVB200_META_PARTNER_BIAS];
                    const mixed mass2 = rcpMass(q.w);
                    const mixed redMass = mass * mass2 * rcpMass(mass + mass2);
                    const mixed rx = v.x - q.x, ry = v.y - q.y, rz = v.z - q.z;
                    pairRel[it] = (rx * rx + ry * ry + rz * rz) * redMass;
                }
            }
        }
        // ---- molecular centre-of-mass velocities (drudeNoseHoover.cu:11-30), COM_LANES lanes per molecule ----
        // A tile of the ionic-liquid box holds ~27 molecules: 108 of the 256 threads, i.e. the block's first warps, carry
        // the whole molecule phase (~280 instructions a tile against ~140 for the particle phase) while the others run
        // ahead by what the ring allows and then spin on a `full` barrier.  No barrier separates the tiles here, so the
        // molecule lanes ROTATE by half a block with every tile a block takes and each warp is the heavy one every other
        // tile.  (The block's first tile is unrotated: a system of one tile per block -- every bitwise fused-vs-split
        // test -- sums in pass A's thread order.  With several tiles per block the molecule terms of every other tile reach
        // the block's sum through other threads than in pass A: the group energies then agree to rounding, <= 1e-15
        // relative; the scale factors are exp() of something 1e-5 small and do not see it -- split and fused trajectories
        // of a 1M-particle box were still bit-identical after three steps.)
        if (nMol > 0) {
            const int rtid = (tid + rot) & (CTHREADS - 1);
            const int grp = rtid / COM_LANES, sub = rtid % COM_LANES;
            for (int jb = 0; jb < nMol; jb += CTHREADS / COM_LANES) {
                if (jb + (rtid & ~31) / COM_LANES >= nMol)
                    break;      // warp-uniform (see passAPhase23): the warps behind the tile's last molecule go on
                const int j = jb + grp;
                const bool active = j < nMol;
                mixed sx = 0, sy = 0, sz = 0, comMass = 0;
                uint32_t info = 0;
                int mol = 0;
                if (active) {
                    info = (uint32_t) st.molInfo[ml0 + j];
                    mol = molFirst >= 0 ? molFirst + j : p.tileMolList[m0 + j];
                    if (!MOLINFO_SCATTERED(info)) {
                        const int first = MOLINFO_FIRST(info), cnt = MOLINFO_COUNT(info);
                        for (int k = first + sub; k < first + cnt; k += COM_LANES) {
                            const mixed4 q = sv[k];
                            const mixed mass = q.w != 0 ? rcpMass(q.w) : (mixed) 0;
                            sx += q.x * mass; sy += q.y * mass; sz += q.z * mass;
                            comMass += mass;
                        }
                    } else if (sub == 0) {
                        const int cnt = p.particlesInMolecules[2 * mol], start = p.particlesInMolecules[2 * mol + 1];
                        for (int k = 0; k < cnt; k++) {
                            const int loc = p.sortedByMol[start + k] - t0;
                            if (loc < 0 || loc >= t1 - t0) continue;
                            const mixed4 q = sv[loc];
                            const mixed mass = q.w != 0 ? rcpMass(q.w) : (mixed) 0;
                            sx += q.x * mass; sy += q.y * mass; sz += q.z * mass;
                            comMass += mass;
                        }
                    }
                }
#pragma unroll
                for (int off = COM_LANES / 2; off > 0; off >>= 1) {
                    sx += __shfl_xor_sync(0xffffffffu, sx, off);
                    sy += __shfl_xor_sync(0xffffffffu, sy, off);
                    sz += __shfl_xor_sync(0xffffffffu, sz, off);
                    comMass += __shfl_xor_sync(0xffffffffu, comMass, off);
                }
                if (active && sub == 0 && MOLINFO_FRAGMENT(info)) {
                    double *fp = p.fragPartials + 5 * (size_t) p.tileMolFrag[m0 + j];
                    fp[0] = (double) sx; fp[1] = (double) sy; fp[2] = (double) sz; fp[3] = (double) comMass; fp[4] = 0.0;
                } else if (active && sub == 0) {
                    mixed4 V;
                    V.w = vv_recip(comMass);
                    V.x = sx * V.w; V.y = sy * V.w; V.z = sz * V.w;
                    st_stream(comV + mol, V);
                    const mixed mv2 = (V.x * V.x + V.y * V.y + V.z * V.z) * comMass;
                    acc[1] += mv2;
                    acc[0] -= mv2;
                }
            }
        }
        mbarArrive(empty + s);      // this thread is done reading the stage
        if (++s == stages) { s = 0; phase ^= 1; }
        rot ^= ROT_RED;
        // the pair terms join the sums AFTER the molecule terms: the order pass A adds them in (bit-identical sums)
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            acc[2] += pairRel[it];
            acc[0] -= pairRel[it];
        }
    }

    if (!blockReduceAndTicket<3>(p, sm, acc, tid))
        return;
    if (p.fuseNHC) {
        lastBlockFinish<MODE, 3>(p, sm, false, tid, &nhcS, true, 0, 1ull, p.flagSync != 0);
        consumerBarrier();
        nhcStore(p.nhc, &nhcS, tid);
    } else {
        lastBlockFinish<MODE, 3>(p, sm, false, tid);
    }
}

// ------------------------------------------------------------------------------------------------
// pass B: thermostat scaling + bias remove/restore + drifts + position write + hard wall
// ------------------------------------------------------------------------------------------------
template <int MODE, int VARIANT, bool EXTRA> struct StageB {
    static constexpr bool POS = VARIANT != VAR_SCALE_ONLY && VARIANT != VAR_SCALE_DELTA;
    static constexpr bool POSQ = POS || EXTRA;
    static constexpr bool CORR = POS && Prec<MODE>::kMixed;
    static constexpr bool FORCE = VARIANT == VAR_VV_FIRST;
    typename Prec<MODE>::mixed4 velm[PADT];
    typename Prec<MODE>::mixed4 comV[MAXMOL];
    uint32_t meta[PADT];
    int32_t desc[8];
    typename Prec<MODE>::real4 posq[POSQ ? PADT : 1];
    typename Prec<MODE>::real4 corr[CORR ? PADT : 1];
    long long f[3][FORCE ? PADT : 2];
    typename Prec<MODE>::mixed cbar[EXTRA ? MAXMOL + 8 : 2];
    // posDelta (and, middle scheme, oldDelta) after OpenMM's position constraints
    typename Prec<MODE>::mixed4 pd[VARIANT == VAR_FINISH || VARIANT == VAR_VV_POSITIONS ? PADT : 1];
    typename Prec<MODE>::mixed4 od[VARIANT == VAR_FINISH ? PADT : 1];
};

template <int MODE, int VARIANT, bool EXTRA> constexpr size_t smemBytesB(int stages) {
    return roundUp128(sizeof(StageB<MODE, VARIANT, EXTRA>)) * stages + 16 * stages + 128;
}

// thermostat scaling of one Drude pair (drudeNoseHoover.cu:164-208).  v* are COM-normalised (and bias-free)
// velocities, m1f/m2f the mass fractions m_k/(m1+m2); returns the new absolute velocities.
template <class mixed>
__device__ __forceinline__ void scalePair(const mixed v1[3], const mixed v2[3], mixed m1f, mixed m2f, const mixed V[3],
                                          mixed sA, mixed sC, mixed sD, mixed out1[3], mixed out2[3]) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
        mixed cm = v1[d] * m1f + v2[d] * m2f;
        mixed rel = v2[d] - v1[d];
        cm = sA * cm;
        rel = sD * rel;
        out1[d] = cm - rel * m2f + sC * V[d];
        out2[d] = cm + rel * m1f + sC * V[d];
    }
}

template <int MODE>
__device__ __forceinline__ void splitPos(typename Prec<MODE>::mixed x, typename Prec<MODE>::real &hi,
                                         typename Prec<MODE>::real &lo) {
    typedef typename Prec<MODE>::real real;
    hi = (real) x;
    lo = (real) (x - (typename Prec<MODE>::mixed) hi);
}

// Everything pass B needs besides the stage: step constants and this step's thermostat factors.
template <int MODE> struct BCtx {
    typedef typename Prec<MODE>::real real;
    typedef typename Prec<MODE>::mixed mixed;
    mixed stepSize, halfdt, invStepSize, fscaleVV, sA, sC, sD, Vb, maxD, hwScale;
    real efscale, accel, invBoxZ;
    bool cosine, useCOM;
    bool writeAllVel;   // resident kernel: pass A's kick is still in shared memory, every massive particle is written
};

// `nhc` values are read through L2 (__ldcg): in the resident kernel another block wrote them during this launch
// `fac`: the factors as pass B's producer received them (hand-over); nullptr: from the thermostat state
// conservative pre-test of the hard wall: below this squared distance `rInv*maxD < 1` cannot hold (KParams::maxD2safe*)
template <class real> __device__ __forceinline__ real hardwallPretest(const KParams &p);
template <> __device__ __forceinline__ float hardwallPretest<float>(const KParams &p) { return p.maxD2safeF; }
template <> __device__ __forceinline__ double hardwallPretest<double>(const KParams &p) { return p.maxD2safeD; }

template <int MODE> __device__ __forceinline__ BCtx<MODE> makeBCtx(const KParams &p, bool extra, const double *fac = nullptr) {
    typedef typename Prec<MODE>::real real;
    typedef typename Prec<MODE>::mixed mixed;
    BCtx<MODE> c;
    c.cosine = extra && p.cosine;
    c.useCOM = p.useCOM;
    c.writeAllVel = false;
    c.stepSize = (mixed) p.dt;
    c.halfdt = 0.5f * c.stepSize;                       // middle.cu:33,51
    c.invStepSize = (mixed) (1.0 / c.stepSize);         // velocityVerlet.cu:40
    c.fscaleVV = (mixed) (0.5 * p.dt / (double) 0x100000000);
    if (fac) {
        c.sA = (mixed) fac[0];
        c.sC = (mixed) fac[1];
        c.sD = (mixed) fac[2];
        c.Vb = c.cosine ? (mixed) fac[3] : (mixed) 0;
    } else {
        c.sA = (mixed) __ldcg(&p.nhc->vscale[0]);
        c.sC = (mixed) __ldcg(&p.nhc->vscale[1]);
        c.sD = (mixed) __ldcg(&p.nhc->vscale[2]);
        c.Vb = c.cosine ? (mixed) __ldcg(&p.nhc->vBias) : (mixed) 0;
    }
    c.maxD = (mixed) p.maxDrudeDistance;
    c.hwScale = (mixed) p.hardwallScale;
    c.efscale = (real) p.efscale;
    c.accel = (real) p.accel;
    c.invBoxZ = (real) p.invBoxZ;
    return c;
}

// One tile of pass B from a filled stage (used by the streaming kernel below and the resident kernel).
// CT = consumer threads working on the tile (each takes ceil(tile / CT) particles)
template <int MODE, int VARIANT, bool EXTRA, int CT, class Stage>
__device__ __forceinline__ void passBTile(const KParams &p, const BCtx<MODE> &cx, Stage &st, const int tid) {
    constexpr int ITEMS_B = (VVB200_TILE_CAP + CT - 1) / CT;
    typedef Prec<MODE> P;
    typedef typename P::real real;
    typedef typename P::mixed mixed;
    typedef typename P::real4 real4;
    typedef typename P::mixed4 mixed4;
    typedef typename P::real3 real3;
    constexpr bool POS = Stage::POS;
    mixed4 *velm = reinterpret_cast<mixed4 *>(p.velm);
    real4 *posq = reinterpret_cast<real4 *>(p.posq);
    real4 *corr = reinterpret_cast<real4 *>(p.corr);
    const real3 *ldForce = reinterpret_cast<const real3 *>(p.ldForce);
    const mixed stepSize = cx.stepSize, halfdt = cx.halfdt, invStepSize = cx.invStepSize, fscaleVV = cx.fscaleVV;
    const mixed sA = cx.sA, sC = cx.sC, sD = cx.sD, Vb = cx.Vb, maxD = cx.maxD, hwScale = cx.hwScale;
    // (the hard wall's pre-test threshold is read from the parameter block where it is used: a constant-bank operand, not
    // one more loop-invariant register -- the velocity-Verlet variant spilled it and reloaded it once per tile and warp)
    const real efscale = cx.efscale, accel = cx.accel, invBoxZ = cx.invBoxZ;
    const bool cosine = cx.cosine, useCOM = cx.useCOM;
    (void) posq; (void) corr; (void) ldForce; (void) invStepSize; (void) fscaleVV; (void) efscale; (void) accel;
    const int t0 = st.desc[0], t1 = st.desc[1], cbOff = st.desc[5];
    const int sl0 = t0 - (t0 & ~3);

#pragma unroll
    for (int it = 0; it < ITEMS_B; it++) {
        const int loc = it * CT + tid;
        const int idx = t0 + loc, sl = sl0 + loc;
        if (idx >= t1) continue;
        const uint32_t mw = st.meta[sl];
        const mixed4 vel = st.velm[sl];
        real4 pq;
        pq.x = pq.y = pq.z = pq.w = 0;
        mixed xs[3] = {0, 0, 0};
        real4 cs;
        cs.x = cs.y = cs.z = cs.w = 0;
        if (Stage::POSQ) {
            pq = st.posq[sl];
            xs[0] = pq.x; xs[1] = pq.y; xs[2] = pq.z;
            if (Stage::CORR) {
                cs = st.corr[sl];
                xs[0] = pq.x + (mixed) cs.x;       // middle.cu:82-84
                xs[1] = pq.y + (mixed) cs.y;
                xs[2] = pq.z + (mixed) cs.z;
            }
        }
        double cphs = 0, cq = 0;
        if (cosine) cphs = cosPhase((double) pq.z, (double) invBoxZ);

        const bool isNH = mw & VVB200_META_NH;
        const uint32_t role = (mw >> VVB200_META_ROLE_SHIFT) & VVB200_META_ROLE_MASK;
        const uint32_t lm = mw & VVB200_META_MOL_MASK;
        const int psl = sl + (int) (mw >> VVB200_META_PARTNER_SHIFT) - VVB200_META_PARTNER_BIAS;
        const bool hasMol = useCOM && lm != VVB200_META_MOL_NONE;
        mixed V[3] = {0, 0, 0};
        mixed cb = 0;
        if (hasMol) {
            const mixed4 Vm = st.comV[lm];
            V[0] = Vm.x; V[1] = Vm.y; V[2] = Vm.z;
            if (cosine) cb = st.cbar[cbOff + lm];
        }
        // this particle ("s") and, for pair roles, its partner ("q")
        mixed vs[3] = {vel.x, vel.y, vel.z};
        const mixed ws = vel.w;
        mixed vq[3] = {0, 0, 0}, wq = 0;
        // the partner's position stays in its raw (posq, posqCorrection) form: it is only converted when the
        // hard wall's pre-test cannot rule the wall out
        real4 pqq, cqr;
        pqq.x = pqq.y = pqq.z = pqq.w = 0;
        cqr.x = cqr.y = cqr.z = cqr.w = 0;
        mixed ds[3] = {0, 0, 0}, dq[3] = {0, 0, 0};      // position increments of this particle / its partner
        if (role != VVB200_ROLE_NONE) {
            const mixed4 v2 = st.velm[psl];
            vq[0] = v2.x; vq[1] = v2.y; vq[2] = v2.z; wq = v2.w;
            if (Stage::POSQ) {
                pqq = st.posq[psl];
                if (Stage::CORR) cqr = st.corr[psl];
                if (cosine) cq = cosPhase((double) pqq.z, (double) invBoxZ);
            }
        }
        // velocities entering the drift as "pre-thermostat" values (middle.cu:33-41)
        const mixed vs0[3] = {vs[0], vs[1], vs[2]};
        const mixed vq0[3] = {vq[0], vq[1], vq[2]};
        bool writeVel = cx.writeAllVel && ws != 0;

        if (VARIANT == VAR_FINISH || VARIANT == VAR_VV_POSITIONS) {
            // no thermostat here: it ran before OpenMM's position constraints
        } else if (isNH) {
            // removePeriodicVelocityBias (cosineAccelerate.cu:63-71)
            if (cosine) { vs[0] -= Vb * cphs; vq[0] -= Vb * cq; }
            // bias-removed molecular velocity: V' = V - Vb*cbar e_x
            mixed Vn[3] = {V[0], V[1], V[2]};
            if (cosine && hasMol) Vn[0] = V[0] - Vb * cb;
            if (hasMol) {   // normalizeVelocities, drudeNoseHoover.cu:42-48
#pragma unroll
                for (int d = 0; d < 3; d++) { vs[d] -= Vn[d]; vq[d] -= Vn[d]; }
            }
            if (role == VVB200_ROLE_NONE) {
                if (ws != 0) {
#pragma unroll
                    for (int d = 0; d < 3; d++) vs[d] = sA * vs[d] + sC * Vn[d];   // drudeNoseHoover.cu:172-176
                }
                writeVel = cosine || hasMol || ws != 0;
            } else {
                // mass fractions m_k/(m1+m2) = w_other/(w1+w2): one division per pair member
                const mixed invW = vv_recip(ws + wq);
                const mixed fs = wq * invW, fq = ws * invW;
                mixed o1[3], o2[3];
                if (role == VVB200_ROLE_DRUDE) {
                    scalePair<mixed>(vs, vq, fs, fq, Vn, sA, sC, sD, o1, o2);
#pragma unroll
                    for (int d = 0; d < 3; d++) { vs[d] = o1[d]; vq[d] = o2[d]; }
                } else {
                    scalePair<mixed>(vq, vs, fq, fs, Vn, sA, sC, sD, o1, o2);
#pragma unroll
                    for (int d = 0; d < 3; d++) { vq[d] = o1[d]; vs[d] = o2[d]; }
                }
                writeVel = true;
            }
            if (cosine) { vs[0] += Vb * cphs; vq[0] += Vb * cq; }   // restorePeriodicVelocityBias
        } else if (cosine) {
            // non-thermostatted atoms still see remove then restore (cosineAccelerate.cu:63-84)
            vs[0] -= Vb * cphs; vs[0] += Vb * cphs;
            vq[0] -= Vb * cq; vq[0] += Vb * cq;
            writeVel = true;
        }

        if (VARIANT == VAR_SCALE_ONLY) {
            if (writeVel) {
                mixed4 o; o.x = vs[0]; o.y = vs[1]; o.z = vs[2]; o.w = ws;
                st_stream(velm + idx, o);
            }
            continue;
        }
        if (VARIANT == VAR_SCALE_DELTA) {
            // integrateMiddlePos1 + Pos2 (middle.cu:33-41, 51-59): posDelta = oldDelta = dt/2 v0 + dt/2 v', for
            // OpenMM's position constraints to work on before vvb200_middle_finish
            if (writeVel) {
                mixed4 o; o.x = vs[0]; o.y = vs[1]; o.z = vs[2]; o.w = ws;
                st_stream(velm + idx, o);
            }
            if (ws != 0) {
                mixed4 d;
                d.x = halfdt * vs0[0]; d.x += halfdt * vs[0];
                d.y = halfdt * vs0[1]; d.y += halfdt * vs[1];
                d.z = halfdt * vs0[2]; d.z += halfdt * vs[2];
                d.w = 0;
                st_stream(reinterpret_cast<mixed4 *>(p.posDelta) + idx, d);
                st_stream(reinterpret_cast<mixed4 *>(p.oldDelta) + idx, d);
            }
            continue;
        }

        bool writePos = false;
        if constexpr (VARIANT == VAR_FINISH) {
            // integrateMiddlePos3 (middle.cu:66-100): v += (posDelta - oldDelta) / dt, x += posDelta, for this particle
            // and (redundantly, for the hard wall) its partner
            const mixed invDt = 1 / stepSize;
            if (ws != 0) {
                const mixed4 d = st.pd[sl], o = st.od[sl];
                vs[0] += (d.x - o.x) * invDt; vs[1] += (d.y - o.y) * invDt; vs[2] += (d.z - o.z) * invDt;
                ds[0] = d.x; ds[1] = d.y; ds[2] = d.z;
#pragma unroll
                for (int k = 0; k < 3; k++) xs[k] += ds[k];
                writePos = writeVel = true;
            }
            if (role != VVB200_ROLE_NONE && wq != 0) {
                const mixed4 d = st.pd[psl], o = st.od[psl];
                vq[0] += (d.x - o.x) * invDt; vq[1] += (d.y - o.y) * invDt; vq[2] += (d.z - o.z) * invDt;
                dq[0] = d.x; dq[1] = d.y; dq[2] = d.z;
            }
        } else if constexpr (VARIANT == VAR_VV_POSITIONS) {
            // velocityVerletIntegratePositions (velocityVerlet.cu:35-68): x += posDelta, v = posDelta / dt
            if (ws != 0) {
                const mixed4 d = st.pd[sl];
                ds[0] = d.x; ds[1] = d.y; ds[2] = d.z;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    vs[k] = (mixed) (invStepSize * ds[k]);
                    xs[k] += ds[k];
                }
                writePos = writeVel = true;
            }
            if (role != VVB200_ROLE_NONE && wq != 0) {
                const mixed4 d = st.pd[psl];
                dq[0] = d.x; dq[1] = d.y; dq[2] = d.z;
#pragma unroll
                for (int k = 0; k < 3; k++) vq[k] = (mixed) (invStepSize * dq[k]);
            }
        } else if (VARIANT == VAR_VV_FIRST) {
            // half kick with the forces of the current positions (velocityVerlet.cu:14-27), for this
            // particle and (redundantly) its partner
            for (int who = 0; who < 2; who++) {
                if (who == 1 && role == VVB200_ROLE_NONE) break;
                const int js = who == 0 ? sl : psl;
                mixed *v = who == 0 ? vs : vq;
                const mixed w = who == 0 ? ws : wq;
                if (w == 0) continue;
                real ex = 0, ey = 0, ez = 0;
                if (EXTRA && p.extraForces) {
                    const uint32_t mj = who == 0 ? mw : st.meta[js];
                    if (p.hasLD && (mj & VVB200_META_LD)) {
                        const real3 f = ldForce[p.ldSlot[t0 + js - sl0]];
                        ex = f.x; ey = f.y; ez = f.z;
                    }
                    if (p.hasField) {
                        const int cnt = (mj >> VVB200_META_ELEC_SHIFT) & VVB200_META_ELEC_MASK;
                        const real q = who == 0 ? pq.w : pqq.w;
                        for (int c = 0; c < cnt; c++) ez += efscale * q;
                    }
                    if (cosine) {
                        const double c = who == 0 ? cphs : cq;
                        ex = (real) (ex + accel * c * vv_recip(w));
                    }
                }
                const long long fx = st.f[0][Stage::FORCE ? js : 0], fy = st.f[1][Stage::FORCE ? js : 0],
                                fz = st.f[2][Stage::FORCE ? js : 0];
                v[0] += 0.5 * stepSize * w * ex + fscaleVV * w * fx;
                v[1] += 0.5 * stepSize * w * ey + fscaleVV * w * fy;
                v[2] += 0.5 * stepSize * w * ez + fscaleVV * w * fz;
            }
            // posDelta = dt*v ; x += posDelta ; v = posDelta/dt  (velocityVerlet.cu:25,52-58)
            if (ws != 0) {
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    ds[d] = stepSize * vs[d];
                    xs[d] += ds[d];
                    vs[d] = (mixed) (invStepSize * ds[d]);
                }
                writePos = writeVel = true;
            }
            if (role != VVB200_ROLE_NONE && wq != 0) {
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    dq[d] = stepSize * vq[d];
                    vq[d] = (mixed) (invStepSize * dq[d]);
                }
            }
        } else {
            // middle scheme without constraints: posDelta = oldDelta = halfdt*v0 + halfdt*v', so
            // integrateMiddlePos3 leaves v' unchanged and moves x by posDelta (middle.cu:33-98)
            if (ws != 0) {
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    ds[d] = halfdt * vs0[d];
                    ds[d] += halfdt * vs[d];
                    xs[d] += ds[d];
                }
                writePos = writeVel = true;
            }
            if (role != VVB200_ROLE_NONE && wq != 0) {
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    dq[d] = halfdt * vq0[d];
                    dq[d] += halfdt * vq[d];
                }
            }
        }

        // ---- Drude hard wall (middle.cu:114-220), evaluated by both members of the pair ----------
        if (p.hardwall && role != VVB200_ROLE_NONE) {
            // Pre-test in `real` arithmetic on the raw operands: the new separation is (posq_s - posq_q) +
            // (corr_s - corr_q) + (ds - dq); the first difference is exact or off by < 4e-9 nm (neighbouring floats),
            // so the squared distance is good to ~1e-6 relative and a 1e-4 margin is conservative.  Only a pair
            // that might touch the wall pays for the partner's fp64 position and the reference's sqrt / reciprocal.
            const real sx = (pq.x - pqq.x) + (cs.x - cqr.x) + (real) (ds[0] - dq[0]);
            const real sy = (pq.y - pqq.y) + (cs.y - cqr.y) + (real) (ds[1] - dq[1]);
            const real sz = (pq.z - pqq.z) + (cs.z - cqr.z) + (real) (ds[2] - dq[2]);
            if (!(sx * sx + sy * sy + sz * sz < hardwallPretest<real>(p))) {
                mixed xq[3] = {pqq.x + (mixed) cqr.x, pqq.y + (mixed) cqr.y, pqq.z + (mixed) cqr.z};
#pragma unroll
                for (int d = 0; d < 3; d++) xq[d] += dq[d];
                // the reference re-reads positions from posq (+ posqCorrection): apply the same rounding
                if (P::kMixed) {
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        real hi, lo;
                        if (ws != 0) { splitPos<MODE>(xs[d], hi, lo); xs[d] = hi + (mixed) lo; }
                        if (wq != 0) { splitPos<MODE>(xq[d], hi, lo); xq[d] = hi + (mixed) lo; }
                    }
                }
                const bool selfIsDrude = role == VVB200_ROLE_DRUDE;
                mixed *pos1 = selfIsDrude ? xs : xq, *pos2 = selfIsDrude ? xq : xs;
                const mixed dx = pos1[0] - pos2[0], dy = pos1[1] - pos2[1], dz = pos1[2] - pos2[2];
                const mixed d2 = dx * dx + dy * dy + dz * dz;
                mixed *vel1 = selfIsDrude ? vs : vq, *vel2 = selfIsDrude ? vq : vs;
                const mixed w1 = selfIsDrude ? ws : wq, w2 = selfIsDrude ? wq : ws;
                const mixed r = vv_sqrt<MODE, mixed>(d2);
                const mixed rInv = vv_recip(r);
                if (rInv * maxD < 1) {
                    const mixed bond[3] = {dx * rInv, dy * rInv, dz * rInv};
                    const mixed mass1 = vv_recip(w1), mass2 = vv_recip(w2);
                    const mixed deltaR = r - maxD;
                    mixed deltaT = stepSize;
                    mixed dotvr1 = vel1[0] * bond[0] + vel1[1] * bond[1] + vel1[2] * bond[2];
                    mixed vp1[3];
#pragma unroll
                    for (int d = 0; d < 3; d++) vp1[d] = vel1[d] - bond[d] * dotvr1;
                    if (w2 == 0) {
                        if (dotvr1 != 0) deltaT = deltaR / fabs(dotvr1);
                        if (deltaT > stepSize) deltaT = stepSize;
                        dotvr1 = -dotvr1 * hwScale / (fabs(dotvr1) * vv_sqrt<MODE, mixed>(mass1));
                        const mixed dr = -deltaR + deltaT * dotvr1;
#pragma unroll
                        for (int d = 0; d < 3; d++) {
                            pos1[d] += bond[d] * dr;
                            vel1[d] = vp1[d] + bond[d] * dotvr1;
                        }
                    } else {
                        const mixed invTotalMass = vv_recip(mass1 + mass2);
                        mixed dotvr2 = vel2[0] * bond[0] + vel2[1] * bond[1] + vel2[2] * bond[2];
                        mixed vp2[3];
#pragma unroll
                        for (int d = 0; d < 3; d++) vp2[d] = vel2[d] - bond[d] * dotvr2;
                        const mixed vbCMass = (mass1 * dotvr1 + mass2 * dotvr2) * invTotalMass;
                        dotvr1 -= vbCMass;
                        dotvr2 -= vbCMass;
                        if (dotvr1 != dotvr2) deltaT = deltaR / fabs(dotvr1 - dotvr2);
                        if (deltaT > stepSize) deltaT = stepSize;
                        const mixed vBond = hwScale / vv_sqrt<MODE, mixed>(mass1);
                        dotvr1 = -dotvr1 * vBond * mass2 * invTotalMass / fabs(dotvr1);
                        dotvr2 = -dotvr2 * vBond * mass1 * invTotalMass / fabs(dotvr2);
                        const mixed dr1 = -deltaR * mass2 * invTotalMass + deltaT * dotvr1;
                        const mixed dr2 = deltaR * mass1 * invTotalMass + deltaT * dotvr2;
                        dotvr1 += vbCMass;
                        dotvr2 += vbCMass;
#pragma unroll
                        for (int d = 0; d < 3; d++) {
                            pos1[d] += bond[d] * dr1;
                            pos2[d] += bond[d] * dr2;
                            vel1[d] = vp1[d] + bond[d] * dotvr1;
                            vel2[d] = vp2[d] + bond[d] * dotvr2;
                        }
                    }
                    // the reference writes the touched members unconditionally (middle.cu:166-172, 204-219)
                    if (selfIsDrude || w2 != 0) writePos = writeVel = true;
                }
            }
        }

        if (writeVel) {
            mixed4 o; o.x = vs[0]; o.y = vs[1]; o.z = vs[2]; o.w = ws;
            st_stream(velm + idx, o);
        }
        real4 o = pq, oc = cs;      // this particle's position as the arrays hold it after this step
        if (POS && writePos) {
            if (P::kMixed) {
                splitPos<MODE>(xs[0], o.x, oc.x);
                splitPos<MODE>(xs[1], o.y, oc.y);
                splitPos<MODE>(xs[2], o.z, oc.z);
                o.w = pq.w;
                oc.w = 0;
                st_stream(posq + idx, o);
                st_stream(corr + idx, oc);
            } else {
                o.x = (real) xs[0]; o.y = (real) xs[1]; o.z = (real) xs[2]; o.w = pq.w;
                st_stream(posq + idx, o);
            }
        }
        // updateImagePositions (imageCharge.cu:2-27) by the PARENT's thread: x, y copied, z mirrored, from the position
        // just written; the image's charge and correction .w are left alone (three scalar stores each).  The image
        // particle itself is massless and outside the thermostat: its own thread writes nothing.
        if (POS && p.imageFused && (mw & VVB200_META_HAS_IMAGE)) {
            const int img = p.imageOf[idx];
            real *ip = reinterpret_cast<real *>(posq + img);
            ip[0] = o.x;
            ip[1] = o.y;
            if (P::kMixed) {
                real *ic = reinterpret_cast<real *>(corr + img);
                ic[0] = oc.x;
                ic[1] = oc.y;
                mixed z = (mixed) o.z + (mixed) oc.z;
                z = (mixed) p.mirror * 2 - z;
                ip[2] = (real) z;
                ic[2] = (real) (z - (real) z);
            } else {
                ip[2] = 2 * (mixed) p.mirror - o.z;
            }
        }
    }
}

template <int MODE, int VARIANT, bool EXTRA>
__global__ void __launch_bounds__(passBConsumers(VARIANT) + 32, passBMinBlocks(VARIANT)) scale_drift_kernel(const KParams p) {
    constexpr int CTHREADS_B = passBConsumers(VARIANT);
    typedef Prec<MODE> P;
    typedef typename P::real real;
    typedef typename P::mixed mixed;
    typedef typename P::real4 real4;
    typedef typename P::mixed4 mixed4;
    typedef typename P::real3 real3;
    typedef StageB<MODE, VARIANT, EXTRA> Stage;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const int stages = p.stagesB;
    constexpr size_t stageBytes = roundUp128(sizeof(Stage));
    uint64_t *full = reinterpret_cast<uint64_t *>(smemRaw + stageBytes * stages);
    uint64_t *empty = full + stages;
    uint64_t *ready = empty + stages;       // hand-over: opened by the producer once the scale factors are published
    __shared__ int expiredS;
    __shared__ double facS[4];

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < stages; s++) {
            mbarInit(full + s, 1);
            mbarInit(empty + s, CTHREADS_B);
        }
        mbarInit(ready, 4);      // the four lanes that poll the factor records
        fenceBarrierInit();
        expiredS = 0;
    }
    __syncthreads();
    // PDL: everything above overlapped pass A's tail.  With the hand-over word this block goes on as soon as what it
    // needs exists (see "early hand-over"); otherwise it waits for the whole preceding grid.
    const bool handOver = p.flagSync && VARIANT != VAR_FINISH && VARIANT != VAR_VV_POSITIONS;
    if (tid == 0) traceMarkS(1, 0);
    if (!handOver) gridDepWait();

    const bool cosine = EXTRA && p.cosine;
    const bool useCOM = p.useCOM && VARIANT != VAR_FINISH && VARIANT != VAR_VV_POSITIONS;   // these two do not touch the thermostat

    if (tid >= CTHREADS_B) {
        // ===== producer warp (all lanes stay: the gather fallback uses them) =====
        const int lane = tid - CTHREADS_B;
        const mixed4 *comV = reinterpret_cast<const mixed4 *>(p.comV);
        const mixed *comCbar = reinterpret_cast<const mixed *>(p.comCbar);
        if (handOver) {
            // pass A's velocities and centre-of-mass velocities must be visible before the first copy is issued; the
            // copies run in the async proxy, hence the proxy fence after the acquire
            if (lane == 0) waitFlagAtLeast(p.syncFlag, 1u);
            __syncwarp();
            asm volatile("fence.proxy.async;" ::: "memory");
            if (lane == 0) traceMarkS(1, 1);
        }
        int s = 0;
        uint32_t phase = 0;
        // hand-over: tiles come from a shared counter.  The ticket for the NEXT round is requested before this round's
        // copies are issued, so the atomic's round trip is never on the producer's critical path (a producer that
        // draws, waits, reads the descriptor, waits, issues needs longer per tile than the consumers do).
        int nextTicket = 0;
        if (handOver && lane == 0) nextTicket = (int) atomicAdd(p.tileTicket, 1u);
        for (int tile = p.tileBegin + blockIdx.x;; tile += gridDim.x) {
            if (handOver) {
                int t = nextTicket;
                if (lane == 0) nextTicket = (int) atomicAdd(p.tileTicket, 1u);
                t = __shfl_sync(0xffffffffu, t, 0);
                // last tile first: what pass A wrote last is what L2 still holds
                tile = p.tileReverse ? p.tileEnd - 1 - t : p.tileBegin + t;
                const bool none = t >= p.tileEnd - p.tileBegin;
                if (none) {
                    // end mark for the consumers: an empty stage whose descriptor says so
                    mbarWait(empty + s, phase ^ 1);
                    if (lane == 0) {
                        reinterpret_cast<Stage *>(smemRaw + stageBytes * s)->desc[0] = -1;
                        mbarArrive(full + s);
                    }
                    break;
                }
            } else if (tile >= p.tileEnd) {
                break;
            }
            const int4 d0 = __ldg(p.tileDesc + 2 * tile), d1 = __ldg(p.tileDesc + 2 * tile + 1);
            mbarWait(empty + s, phase ^ 1);
            Stage &st = *reinterpret_cast<Stage *>(smemRaw + stageBytes * s);
            const int a0 = d0.x & ~3, cnt = ((d0.y + 3) & ~3) - a0;
            const int nMol = useCOM ? d0.w : 0;
            const bool gather = nMol > 0 && d1.x < 0;
            if (gather) {
                // tile molecules are not consecutive ids: gather through the list (any topology)
                for (int j = lane; j < nMol; j += 32) {
                    const int mol = p.tileMolList[d0.z + j];
                    st.comV[j] = comV[mol];
                    if (cosine) st.cbar[j] = comCbar[mol];
                }
                __syncwarp();
            }
            if (lane == 0) {
                const int cb0 = d1.x >= 0 ? d1.x & ~3 : 0;
                const int cbcnt = cosine && nMol > 0 && !gather ? ((d1.x + nMol + 3) & ~3) - cb0 : 0;
                st.desc[0] = d0.x; st.desc[1] = d0.y; st.desc[2] = d0.z; st.desc[3] = d0.w; st.desc[4] = d1.x;
                st.desc[5] = gather ? 0 : d1.x - cb0;   // offset of this tile's first molecule in st.cbar
                uint32_t bytes = cnt * (uint32_t) (sizeof(mixed4) + sizeof(uint32_t));
                if (Stage::POSQ) bytes += cnt * (uint32_t) sizeof(real4);
                if (Stage::CORR) bytes += cnt * (uint32_t) sizeof(real4);
                if (Stage::FORCE) bytes += 3u * cnt * 8u;
                if (VARIANT == VAR_FINISH) bytes += 2u * cnt * (uint32_t) sizeof(mixed4);
                if (VARIANT == VAR_VV_POSITIONS) bytes += cnt * (uint32_t) sizeof(mixed4);
                if (nMol > 0 && !gather) bytes += nMol * (uint32_t) sizeof(mixed4) + cbcnt * (uint32_t) sizeof(mixed);
                mbarArriveExpectTx(full + s, bytes);
                bulkLoad(st.velm, reinterpret_cast<const mixed4 *>(p.velm) + a0, cnt * (uint32_t) sizeof(mixed4), full + s);
                bulkLoad(st.meta, p.slotMeta + a0, cnt * 4u, full + s);
                if (Stage::POSQ) bulkLoad(st.posq, reinterpret_cast<const real4 *>(p.posq) + a0, cnt * (uint32_t) sizeof(real4), full + s);
                if (Stage::CORR) bulkLoad(st.corr, reinterpret_cast<const real4 *>(p.corr) + a0, cnt * (uint32_t) sizeof(real4), full + s);
                if (Stage::FORCE) {
                    bulkLoad(st.f[0], p.force + a0, cnt * 8u, full + s);
                    bulkLoad(st.f[1], p.force + a0 + p.paddedN, cnt * 8u, full + s);
                    bulkLoad(st.f[2], p.force + a0 + 2 * (size_t) p.paddedN, cnt * 8u, full + s);
                }
                if (VARIANT == VAR_VV_POSITIONS)
                    bulkLoad(st.pd, reinterpret_cast<const mixed4 *>(p.posDelta) + a0, cnt * (uint32_t) sizeof(mixed4), full + s);
                if (VARIANT == VAR_FINISH) {
                    bulkLoad(st.pd, reinterpret_cast<const mixed4 *>(p.posDelta) + a0, cnt * (uint32_t) sizeof(mixed4), full + s);
                    bulkLoad(st.od, reinterpret_cast<const mixed4 *>(p.oldDelta) + a0, cnt * (uint32_t) sizeof(mixed4), full + s);
                }
                if (nMol > 0 && !gather) {
                    bulkLoad(st.comV, comV + d1.x, nMol * (uint32_t) sizeof(mixed4), full + s);
                    if (cbcnt) bulkLoad(st.cbar, comCbar + cb0, cbcnt * (uint32_t) sizeof(mixed), full + s);
                }
            }
            if (++s == stages) { s = 0; phase ^= 1; }
        }
        return;
    }

    // ===== consumers =====
    if (handOver) {
        // lanes 0..3 of the first consumer warp each poll one factor record (the producer is busy filling the ring) and
        // hand the value to everybody through shared memory; the other warps sleep on the barrier
        if (tid < 4) {
            double v = 0;
            const long long c0 = clock64();
            while (!factorPoll(p.factorRec + tid, 1ull, &v)) {
                if (clock64() - c0 > 4000000000LL) { expiredS = 1; break; }
                __nanosleep(20);
            }
            facS[tid] = v;
            mbarArrive(ready);      // each writer arrives for its own value (release)
        }
        mbarWait(ready, 0);
    }
    if (tid == 0) traceMarkS(1, 2);
    BCtx<MODE> cx = makeBCtx<MODE>(p, EXTRA, handOver ? facS : nullptr);
    if (handOver && expiredS) {
        const mixed nan = (mixed) __longlong_as_double(0x7ff8000000000000LL);
        cx.sA = cx.sC = cx.sD = nan;
    }
    int s = 0;
    uint32_t phase = 0;
    // tiles this block takes under the static stride; under the hand-over the producer's end mark ends the loop
    const int mine = p.tileEnd - p.tileBegin - (int) blockIdx.x;
    for (int left = handOver ? 0x7fffffff : (mine > 0 ? (mine + (int) gridDim.x - 1) / (int) gridDim.x : 0); left > 0; left--) {
        mbarWait(full + s, phase);
        Stage &st = *reinterpret_cast<Stage *>(smemRaw + stageBytes * s);
#ifdef VVB200_TRACE
        if (tid == 0 && g_vvb200TraceS[(1024 + blockIdx.x) * 8 + 3] < g_vvb200TraceS[(1024 + blockIdx.x) * 8 + 0]) traceMarkS(1, 3);
#endif
        if (handOver && st.desc[0] < 0) break;      // the producer's end mark
        passBTile<MODE, VARIANT, EXTRA, CTHREADS_B>(p, cx, st, tid);
        mbarArrive(empty + s);   // this thread is done reading the stage
        if (++s == stages) { s = 0; phase ^= 1; }
    }
    if (tid == 0) traceMarkS(1, 4);
    if (tid == 0) gridDepLaunch();      // the next kernel in the stream (next step's pass A) may be scheduled
}
