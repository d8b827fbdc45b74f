// vvb200_resident.cuh -- the whole thermostatted step in ONE launch for systems whose state fits on chip
// (included by vvb200_device.cu after vvb200_stream.cuh).
//
// The streaming kernels (vvb200_stream.cuh) are built for systems much larger than L2: two passes over HBM with
// the group-energy reduction and the NH chains between them.  For the systems the reference's examples actually
// run (9k-50k particles, BASELINE configs 1-4) a step costs ~10 us per launch, not bytes: kernel launch, pipeline
// fill, the tail of the last-block reduction.  Here every block keeps its tiles in shared memory for the whole step:
//
//   load tiles (cp.async.bulk, all issued up front)            velm, force, posq, posqCorrection, slot words
//   pass A per tile (passAPhase1/23 of vvb200_stream.cuh)      kicked velocities written back INTO the stage,
//                                                             molecular velocities into the stage's comV
//   per-block partial sums -> arrival ticket -> grid barrier   the last block to arrive sums the partials, runs the part
//                                                             of the chain update that yields the scale factors
//                                                             (nhcCrit) and publishes them as self-validating records
//                                                             tagged with this launch's generation; every block polls
//                                                             the records (bounded).  The rest of the chain update runs
//                                                             in the last block WHILE it does its pass B, on a warp that
//                                                             holds no particles (tiles of <= 224 slots), else after it
//   pass B per tile (passBTile) straight from the stage        one write of velm / posq / posqCorrection
//
// => 1 launch instead of 2, 156 instead of 224 B/particle (mixed), no HBM round trip of the kicked velocities.
// All blocks must be co-resident: the host only launches it with grid <= (occupancy x SMs), see launchResident.
// Capacity with 227 KB of shared memory: 3 tiles/block x 148 blocks = 227k particles (mixed), more in single.
#pragma once

template <int MODE, int KICK, int VARIANT, bool EXTRA> struct StageR {
    static constexpr bool POS = VARIANT != VAR_SCALE_ONLY && VARIANT != VAR_SCALE_DELTA;
    static constexpr bool POSQ = POS || EXTRA;
    static constexpr bool CORR = POS && Prec<MODE>::kMixed;
    static constexpr bool FORCE = KICK != KICK_NONE || VARIANT == VAR_VV_FIRST;
    typename Prec<MODE>::mixed4 velm[PADT];
    typename Prec<MODE>::mixed4 comV[MAXMOL];
    uint32_t meta[PADT];
    int32_t molInfo[MAXMOL + 8];
    int32_t desc[8];
    typename Prec<MODE>::real4 posq[POSQ ? PADT : 1];
    typename Prec<MODE>::real4 corr[CORR ? PADT : 1];
    long long f[3][FORCE ? PADT : 2];
    typename Prec<MODE>::mixed cbar[EXTRA ? MAXMOL + 8 : 2];
    typename Prec<MODE>::mixed4 pd[1], od[1];      // only the streaming finish variant stages posDelta / oldDelta
};

template <int MODE, int KICK, int VARIANT, bool EXTRA> constexpr size_t smemBytesR(int tiles) {
    return roundUp128(sizeof(StageR<MODE, KICK, VARIANT, EXTRA>)) * tiles + roundUp128(sizeof(ScratchA<MODE, EXTRA>)) + 16 * tiles + 128;
}

#ifndef RESIDENT_MINBLOCKS
#define RESIDENT_MINBLOCKS 2
#endif

template <int MODE, int KICK, int VARIANT, bool EXTRA>
__global__ void __launch_bounds__(CTHREADS, RESIDENT_MINBLOCKS) resident_step_kernel(const KParams p) {
    typedef Prec<MODE> P;
    typedef typename P::mixed mixed;
    typedef typename P::real4 real4;
    typedef typename P::mixed4 mixed4;
    typedef StageR<MODE, KICK, VARIANT, EXTRA> Stage;
    typedef ScratchA<MODE, EXTRA> Scratch;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ unsigned int expiredS, genS, helperS, lastS;
    __shared__ double facS[4];
    __shared__ NhcDevice nhcS;      // every block prefetches the thermostat state: any of them may arrive last
    const int T = p.tilesPerBlock;
    constexpr size_t stageBytes = roundUp128(sizeof(Stage));
    Scratch &sm = *reinterpret_cast<Scratch *>(smemRaw + stageBytes * T);
    uint64_t *full = reinterpret_cast<uint64_t *>(smemRaw + stageBytes * T + roundUp128(sizeof(Scratch)));

    const int tid = threadIdx.x;
    const bool cosine = EXTRA && p.cosine;
    const bool useCOM = p.useCOM;
    traceMark(0);

    unsigned int gen0 = 0;      // thread 0: this launch's barrier generation
    if (tid == 0) {
        // first tile's descriptors and the barrier generation are requested before anything that waits
        int4 d0n = __ldg(p.tileDesc + 2 * blockIdx.x), d1n = __ldg(p.tileDesc + 2 * blockIdx.x + 1);
        gen0 = *reinterpret_cast<volatile const unsigned int *>(p.gridGen);   // read before this block can arrive
        for (int j = 0; j < T; j++) mbarInit(full + j, 1);
        fenceBarrierInit();
        expiredS = 0;
        genS = gen0;
        lastS = 0;
        int widest = 0;
        // ---- every tile of this block is requested up front; nothing is ever refilled ----
        for (int j = 0; j < T; j++) {
            const int tile = blockIdx.x + j * gridDim.x;
            if (tile >= p.numTiles) break;
            const int4 d0 = d0n, d1 = d1n;
            if (tile + (int) gridDim.x < p.numTiles && j + 1 < T) {
                d0n = __ldg(p.tileDesc + 2 * (tile + gridDim.x));
                d1n = __ldg(p.tileDesc + 2 * (tile + gridDim.x) + 1);
            }
            Stage &st = *reinterpret_cast<Stage *>(smemRaw + stageBytes * j);
            const int a0 = d0.x & ~3, cnt = ((d0.y + 3) & ~3) - a0;
            const int ma0 = d0.z & ~3, mcnt = useCOM && d0.w > 0 ? ((d0.z + d0.w + 3) & ~3) - ma0 : 0;
            widest = max(widest, d0.y - d0.x);
            st.desc[0] = d0.x; st.desc[1] = d0.y; st.desc[2] = d0.z; st.desc[3] = d0.w; st.desc[4] = d1.x;
            st.desc[5] = 0;                     // st.cbar is indexed by the tile-local molecule id
            uint32_t bytes = cnt * (uint32_t) (sizeof(mixed4) + sizeof(uint32_t)) + mcnt * 4u;
            if (Stage::FORCE) bytes += 3u * cnt * 8u;
            if (Stage::POSQ) bytes += cnt * (uint32_t) sizeof(real4);
            if (Stage::CORR) bytes += cnt * (uint32_t) sizeof(real4);
            mbarArriveExpectTx(full + j, bytes);
            bulkLoad(st.velm, reinterpret_cast<const mixed4 *>(p.velm) + a0, cnt * (uint32_t) sizeof(mixed4), full + j);
            if (Stage::FORCE) {
                bulkLoad(st.f[0], p.force + a0, cnt * 8u, full + j);
                bulkLoad(st.f[1], p.force + a0 + p.paddedN, cnt * 8u, full + j);
                bulkLoad(st.f[2], p.force + a0 + 2 * (size_t) p.paddedN, cnt * 8u, full + j);
            }
            bulkLoad(st.meta, p.slotMeta + a0, cnt * 4u, full + j);
            if (mcnt) bulkLoad(st.molInfo, p.tileMolInfo + ma0, mcnt * 4u, full + j);
            if (Stage::POSQ) bulkLoad(st.posq, reinterpret_cast<const real4 *>(p.posq) + a0, cnt * (uint32_t) sizeof(real4), full + j);
            if (Stage::CORR) bulkLoad(st.corr, reinterpret_cast<const real4 *>(p.corr) + a0, cnt * (uint32_t) sizeof(real4), full + j);
        }
        // the last warp holds no particle of pass B when no tile of this block is wider than CTHREADS - 32 slots (the
        // tiles of the systems this kernel is for are 128-224 slots): it is free for the rest of the chain update
        helperS = widest <= CTHREADS - 32 ? 1u : 0u;
    }
    if (p.doReduce)
        nhcFetch(&nhcS, p.nhc, tid, 32);
    __syncthreads();
    traceMark(1);

    // ===== pass A on the resident tiles =====
    constexpr int NR = EXTRA ? VVB200_NRED : 3;
    mixed acc[NR];
#pragma unroll
    for (int k = 0; k < NR; k++) acc[k] = 0;
    if (KICK != KICK_NONE || p.doReduce) {
        const ACtx<MODE> ca = makeACtx<MODE, KICK>(p, EXTRA);
        int buf = 0;
        for (int j = 0; j < T; j++) {
            const int tile = blockIdx.x + j * gridDim.x;
            if (tile >= p.numTiles) break;
            mbarWait(full + j, 0);
            if (j == 0) traceMark(2);
            Stage &st = *reinterpret_cast<Stage *>(smemRaw + stageBytes * j);
            PublishedA<MODE, EXTRA> &pub = sm.pub[buf];
            const int t0 = st.desc[0], t1 = st.desc[1], m0 = st.desc[2], nMol = useCOM ? st.desc[3] : 0, molFirst = st.desc[4];
            mixed4 vel[ITEMS];
            uint32_t meta[ITEMS];
            passAPhase1<MODE, KICK, EXTRA, true>(p, ca, st, pub, vel, meta, acc, tid);
            consumerBarrier();
            passAPhase23<MODE, EXTRA, true, Stage, KICK>(p, ca, st, pub, vel, meta, acc, tid, t0, t1, m0, nMol, molFirst);
            buf ^= 1;
        }
    }

    // ===== grid barrier: partial sums -> last block: totals + scale factors -> everyone goes on =====
    traceMark(3);
    constexpr int HELPER0 = CTHREADS - 32;      // first thread of the helper warp
    if (p.doReduce) {
        const bool last = blockReduceAndTicket<NR>(p, sm, acc, tid);
        traceMark(4);
        const unsigned long long tag = (unsigned long long) genS + 2ull;      // >= 2: the streaming kernels use 0 and 1
        if (last) {
            // the three chain threads sit in the helper warp when there is one: after nhcCrit they go straight on to the
            // rest of the chain update while the other seven warps do pass B
            if (tid == 0) lastS = 1;
            lastBlockFinish<MODE, NR>(p, sm, cosine, tid, &nhcS, p.fuseNHC != 0, helperS ? HELPER0 : 0, tag, true, true);
        }
        // every block, the last one included, takes the factors from the records (value + tag in one 128-bit word)
        if (p.fuseNHC) {
            if (tid < 4) {
                double v = 0;
                const long long c0 = clock64();
                while (!factorPoll(p.factorRec + tid, tag, &v)) {
                    if (clock64() - c0 > 2000000000LL) {   // ~1 s: blocks were not co-resident; poison, do not hang
                        expiredS = 1;
                        break;
                    }
                    __nanosleep(20);
                }
                if (p.numSplit != 0) __threadfence();      // acquire side of the cut molecules' velocities
                facS[tid] = v;
            }
        } else if (last) {
            __threadfence();
            if (tid == 0) st_release_gpu(p.gridGen, genS + 1u);
        } else if (tid == 0) {
            const long long c0 = clock64();
            while (ld_acquire_gpu(p.gridGen) == genS) {
                if (clock64() - c0 > 2000000000LL) { expiredS = 1; break; }
                __nanosleep(32);
            }
        }
    }
    __syncthreads();
    traceMark(5);
    const bool helperBusy = p.doReduce && p.fuseNHC && lastS && helperS && tid >= HELPER0;   // this warp: chain update, not pass B

    // molecules cut across tiles were finished by the last block: fetch their velocities into the stages
    if (p.numSplit > 0 && useCOM) {
        for (int j = 0; j < T; j++) {
            const int tile = blockIdx.x + j * gridDim.x;
            if (tile >= p.numTiles) break;
            Stage &st = *reinterpret_cast<Stage *>(smemRaw + stageBytes * j);
            const int m0 = st.desc[2], nMol = st.desc[3], molFirst = st.desc[4], ml0 = m0 - (m0 & ~3);
            if (tid < nMol && MOLINFO_FRAGMENT((uint32_t) st.molInfo[ml0 + tid])) {
                const int mol = molFirst >= 0 ? molFirst + tid : p.tileMolList[m0 + tid];
                const double *src = reinterpret_cast<const double *>(reinterpret_cast<const mixed4 *>(p.comV) + mol);
                mixed4 V;
                if (sizeof(mixed) == 8) { V.x = (mixed) __ldcg(src); V.y = (mixed) __ldcg(src + 1); V.z = (mixed) __ldcg(src + 2); V.w = (mixed) __ldcg(src + 3); }
                else { const float *sf = reinterpret_cast<const float *>(src); V.x = (mixed) __ldcg(sf); V.y = (mixed) __ldcg(sf + 1); V.z = (mixed) __ldcg(sf + 2); V.w = (mixed) __ldcg(sf + 3); }
                st.comV[tid] = V;
                if (EXTRA && p.cosine) st.cbar[tid] = __ldcg(reinterpret_cast<const mixed *>(p.comCbar) + mol);
            }
        }
        __syncthreads();
    }

    // ===== pass B from the same stages =====
    BCtx<MODE> cb = makeBCtx<MODE>(p, EXTRA, p.doReduce && p.fuseNHC ? facS : nullptr);
    cb.writeAllVel = KICK != KICK_NONE;
    if (expiredS) {
        const mixed nan = (mixed) __longlong_as_double(0x7ff8000000000000LL);
        cb.sA = cb.sC = cb.sD = nan;
    }
    if (helperBusy) {
        // the last block's helper warp: the second sweep of the chain update, the look-ahead for the next step, the state
        // back to global memory and the generation word -- concurrently with the other warps' pass B
        if (tid - HELPER0 < 3) nhcPost(&nhcS, p.dt, tid - HELPER0);
        __syncwarp();
        for (int i = tid - HELPER0; i < (int) (sizeof(NhcDevice) / 8); i += 32)
            reinterpret_cast<double *>(p.nhc)[i] = reinterpret_cast<const double *>(&nhcS)[i];
        if (tid == HELPER0) *p.gridGen = genS + 1u;      // read again only by the next launch
    } else {
        for (int j = 0; j < T; j++) {
            const int tile = blockIdx.x + j * gridDim.x;
            if (tile >= p.numTiles) break;
            mbarWait(full + j, 0);
            Stage &st = *reinterpret_cast<Stage *>(smemRaw + stageBytes * j);
            passBTile<MODE, VARIANT, EXTRA, CTHREADS>(p, cb, st, tid);
        }
    }
    traceMark(6);
    if (p.doReduce && p.fuseNHC && lastS && !helperS) {
        // no free warp (tiles wider than 224 slots): the rest of the chain update after this block's pass B
        if (tid < 3) nhcPost(&nhcS, p.dt, tid);
        __syncthreads();
        nhcStore(p.nhc, &nhcS, tid);
        if (tid == 0) *p.gridGen = genS + 1u;
    }
}
