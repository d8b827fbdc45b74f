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
//   per-block partial sums -> arrival ticket -> grid barrier   the last block to arrive sums the partials, advances
//                                                             the NH chains and releases a generation word the
//                                                             other blocks wait on (ld.acquire.gpu, bounded)
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
    __shared__ unsigned int expiredS;
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

    // ===== grid barrier: partial sums -> last block: totals + NH chains -> everyone goes on =====
    traceMark(3);
    if (p.doReduce) {
        const bool last = blockReduceAndTicket<NR>(p, sm, acc, tid);
        traceMark(4);
        if (last) {
            // (the chain update is not split here: the last block has tiles of its own to finish, so what it does after
            // the release is on the launch's critical path either way)
            lastBlockFinish<MODE, NR>(p, sm, cosine, tid, &nhcS);
            consumerBarrier();
            nhcStore(p.nhc, &nhcS, tid);      // the advanced state back to global memory
            consumerBarrier();        // the release below is cumulative over what the barrier ordered before it
            if (tid == 0) st_release_gpu(p.gridGen, gen0 + 1u);
        } else if (tid == 0) {
            const long long c0 = clock64();
            while (ld_acquire_gpu(p.gridGen) == gen0) {
                if (clock64() - c0 > 2000000000LL) {   // ~1 s: blocks were not co-resident; poison, do not hang
                    expiredS = 1;
                    break;
                }
                __nanosleep(32);
            }
        }
    }
    __syncthreads();
    traceMark(5);

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
    BCtx<MODE> cb = makeBCtx<MODE>(p, EXTRA);
    cb.writeAllVel = KICK != KICK_NONE;
    if (expiredS) {
        const mixed nan = (mixed) __longlong_as_double(0x7ff8000000000000LL);
        cb.sA = cb.sC = cb.sD = nan;
    }
    for (int j = 0; j < T; j++) {
        const int tile = blockIdx.x + j * gridDim.x;
        if (tile >= p.numTiles) break;
        mbarWait(full + j, 0);
        Stage &st = *reinterpret_cast<Stage *>(smemRaw + stageBytes * j);
        passBTile<MODE, VARIANT, EXTRA, CTHREADS>(p, cb, st, tid);
    }
    traceMark(6);
}
