// k5_huffman_decode.cu -- K5: Huffman DEcoding of a baseline sequential scan on the device (second half).
// Replaces, for the reference's mj_read_jpeg_from_memory (src/image.c:33-118), the entropy decoder that
// jpeg_read_coefficients runs on the host (libjpeg jdhuff.c decode_mcu): 8-bit sequential DCT, Huffman tables as the file
// states them, ONE scan with every component, no restart markers.  Everything else (progressive, arithmetic, multi-scan,
// restart intervals) stays with libjpeg.
//
// The difficulty of decoding in parallel is that a decoder dropped into the middle of the stream does not know its state:
// the bit position of the next code word, which block of the MCU it is in (which tables apply) and which coefficient comes
// next.  Huffman codes SELF-SYNCHRONISE, though: a decoder started in a wrong state falls into step with the right one
// after a few dozen symbols (Klein & Wiseman 2003; Weissenberger & Schmidt, "Massively parallel Huffman decoding on GPUs",
// ICPP 2018, whose scheme this follows).  One CTA per image:
//   1. un-stuff: every 0x00 behind a 0xFF is dropped (parallel count, CTA-wide prefix sum, scatter), the stream is kept as
//      big-endian 32-bit words so that a bit window is one funnel shift;
//   2. the stream is cut into subsequences (one per thread where the image is large enough: at least kMinSubBits bits each).
//      entry[t] = decoder state at the start of subsequence t:
//      exact for t = 0, a guess (block start of the MCU's first block) for the others.  ROUNDS: every subsequence whose
//      entry state changed is decoded (no output) and its exit state becomes the entry state of the next one.  When a round
//      changes nothing, every entry state is exact by induction from t = 0 -- no probabilistic argument is involved, only
//      the NUMBER of rounds depends on how fast the codes synchronise (3 on the bench's 1080p files);
//   3. prefix sum of the blocks each subsequence completes -> index of the block a subsequence starts in;
//   4. every subsequence is decoded once more, now writing its coefficients (de-zigzagged) into the planes; the DC slot
//      receives the DIFFERENCE;
//   5. per component, a prefix sum over the blocks in scan order turns the differences into DC values.
// Blocks of the scan that lie past a component's real width / height (the MCU grid is rounded up) are decoded into the
// planes' padding exactly as libjpeg does, so the planes equal jpeg_read_coefficients' arrays block for block.
// status[i] != 0: not decoded (a code the tables do not contain, a run past coefficient 63, fewer blocks than the frame
// announces) -- the caller lets libjpeg read that image (and report its error).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mjx_internal.cuh"

namespace mjx {

static constexpr int kDecThreads = 1024;
static constexpr int kMinSubBits = 256; // shortest subsequence; larger images: total bits / threads, so that every thread has ONE
                                        // (the entry states settle in as many rounds as a decoder needs subsequences to synchronise:
                                        // the longer the subsequence, the fewer rounds -- 1 024-bit pieces took 9 rounds on 1080p files)
static constexpr int kLook = 9;       // bits of the first-level code lookup

// one Huffman table, ready for decoding (host-built: ITU-T T.81 F.2.2.3 / libjpeg jdhuff.c jpeg_make_d_derived_tbl)
struct DecTable {
    uint16_t look[1 << kLook]; // (length << 8) | symbol for codes of up to kLook bits, 0: longer
    int32_t  maxcode[18];      // largest code of length l (l = 1..16), -1: none; [17] ends every search
    int32_t  valoff[17];       // index of the first symbol of length l minus the smallest code of length l
    uint8_t  vals[256];
};

struct DecParams {
    const mjx_image_desc_t *items;
    const unsigned char    *data;    // stuffed entropy-coded bytes of all images
    const uint64_t         *offsets; // [n] where image i's segment starts in `data`
    const uint32_t         *lengths; // [n] its length in bytes
    int                     n;
    int                     ncomp, blocks_per_mcu, mcus_per_row, mcu_rows, nblk;
    int                     h[MJX_MAX_COMPONENTS], v[MJX_MAX_COMPONENTS];
    int                     dc_tbl[MJX_MAX_COMPONENTS], ac_tbl[MJX_MAX_COMPONENTS];
    signed char             bcomp[16], bidx[16], byoff[16], bxoff[16]; // per block of the MCU: component, index in it, row / column offset
    const DecTable         *tables; // [8]: 0..3 DC, 4..7 AC
    uint32_t               *words;  // [n][words_stride] un-stuffed stream
    size_t                  words_stride;
    uint2                  *entry;  // [n][sub_stride] state at the start of a subsequence: x = bit position, y = block-in-MCU | z << 8
    uint2                  *exits;  // [n][sub_stride]
    uint32_t               *cnt;    // [n][sub_stride] blocks completed inside the subsequence, then their exclusive prefix sum
    unsigned char          *dirty;  // [n][2][sub_stride] entry state changed: decode again (this round / the next)
    size_t                  sub_stride;
    uint32_t               *status; // [n]
    uint32_t               *rounds; // [n] (diagnostics)
};

// zigzag position -> natural index (libjpeg jutils.c jpeg_natural_order; ITU-T T.81 figure A.6)
__device__ __constant__ unsigned char c_natural[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                                        41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                                        30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// the 32 bits of the stream that start at bit p
__device__ __forceinline__ uint32_t peek32(const uint32_t *words, uint32_t p) {
    const uint32_t i = p >> 5, s = p & 31u;
    return __funnelshift_l(words[i + 1], words[i], s);
}

// value of an s-bit field (T.81 F.2.2.1 EXTEND; libjpeg HUFF_EXTEND)
__device__ __forceinline__ int extend(uint32_t bits, int s) { return (int)bits < (1 << (s - 1)) ? (int)bits - (1 << s) + 1 : (int)bits; }

// CTA-wide exclusive prefix sum of one value per thread (kDecThreads threads); *total receives the sum
__device__ __forceinline__ unsigned long long cta_exclusive_scan(unsigned long long x, unsigned long long *s_warp, unsigned long long *total) {
    const int          lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = x;
#pragma unroll
    for(int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
        if(lane >= o) inc += y;
    }
    __syncthreads(); // s_warp may still be read from the previous call
    if(lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if(warp == 0) {
        unsigned long long y = s_warp[lane];
#pragma unroll
        for(int o = 1; o < 32; o <<= 1) {
            const unsigned long long z = __shfl_up_sync(0xffffffffu, y, o);
            if(lane >= o) y += z;
        }
        s_warp[lane] = y;
    }
    __syncthreads();
    *total = s_warp[kDecThreads / 32 - 1];
    return inc - x + (warp > 0 ? s_warp[warp - 1] : 0ull);
}

struct DecState {
    uint32_t p; // bit position of the next code word
    int      b; // block of the MCU (index into bcomp / bidx)
    int      z; // next coefficient in zigzag order, 0: the DC code comes next
    int      bad;
};
// (`bad` is NOT part of the state that travels between subsequences: a decoder in a wrong state meets invalid codes all the
// time, and carrying the flag along would keep it from ever agreeing with the right one.  The write pass, which starts every
// subsequence in its exact state, raises it afresh.)
__device__ __forceinline__ uint2 pack_state(const DecState &s) { return make_uint2(s.p, (uint32_t)s.b | ((uint32_t)s.z << 8)); }
__device__ __forceinline__ DecState unpack_state(uint2 u) {
    DecState s;
    s.p = u.x, s.b = (int)(u.y & 0xffu), s.z = (int)((u.y >> 8) & 0xffu), s.bad = 0;
    return s;
}

// Bit reader: the next 33..64 bits of the stream left-aligned in a 64-bit register, refilled one 32-bit word at a time -- one
// load per 32 bits consumed instead of two per field looked at (the fields average 6 bits).
struct BitReader {
    const uint32_t    *words;
    unsigned long long buf; // valid bits at the top
    int                cnt; // how many (33..64 between fields)
    uint32_t           wi;  // index of `nxt`
    uint32_t           nxt; // the word the next refill will use, loaded a refill ahead so that nobody waits for it
    __device__ __forceinline__ void start(const uint32_t *w, uint32_t p) {
        words = w;
        wi = p >> 5;
        buf = ((unsigned long long)w[wi] << 32) | w[wi + 1];
        buf <<= (p & 31u);
        cnt = 64 - (int)(p & 31u);
        wi += 2;
        nxt = w[wi];
    }
    __device__ __forceinline__ uint32_t peek32() const { return (uint32_t)(buf >> 32); }
    __device__ __forceinline__ void     skip(int n) {
        buf <<= n;
        cnt -= n;
        if(cnt <= 32) {
            buf |= (unsigned long long)nxt << (32 - cnt);
            cnt += 32;
            nxt = words[++wi];
        }
    }
    __device__ __forceinline__ uint32_t pos() const { return wi * 32u - (uint32_t)cnt; }
};

// one code word of table t: symbol, and the reader moves behind it.  A bit pattern that is no code counts as a 16-bit code for
// symbol 0 and raises bad (a decoder in a wrong state may meet one; the right one never does in a valid file).
__device__ __forceinline__ int decode_symbol(const DecTable &t, BitReader &br, int &bad) {
    const uint32_t v = br.peek32();
    const uint32_t e = t.look[v >> (32 - kLook)];
    if(e) {
        br.skip((int)(e >> 8));
        return (int)(e & 0xffu);
    }
    int l = kLook + 1;
    int code = (int)(v >> (32 - l));
    while(code > t.maxcode[l]) {
        l++;
        code = (int)(v >> (32 - l));
    }
    if(l > 16) {
        bad = 1;
        br.skip(16);
        return 0;
    }
    br.skip(l);
    return (int)t.vals[(code + t.valoff[l]) & 0xff];
}

// Decode from state s to the end of the subsequence [.., end) (the code word that crosses `end` is finished).  kWrite: the
// coefficients go to the planes; blk = index of the block s lies in.  Returns the number of blocks completed.
template <bool kWrite>
__device__ __forceinline__ uint32_t decode_run(const DecParams &p, const mjx_image_desc_t &im, const DecTable *tab, const uint32_t *words, DecState &s, uint32_t end,
                                               uint32_t blk) {
    uint32_t done = 0;
    int16_t *dst = nullptr;
    // where block `blk` lies: MCU (mrow, mcol), block bi of it -- divisions once per subsequence, then counted along
    int  bi = 0, mrow = 0, mcol = 0;
    auto locate = [&]() -> int16_t * {
        const int c = p.bcomp[bi];
        const int row = mrow * p.v[c] + p.byoff[bi], col = mcol * p.h[c] + p.bxoff[bi];
        if(row >= im.rows[c] || col >= im.stride_blocks[c]) return nullptr;
        return reinterpret_cast<int16_t *>(im.plane[c]) + ((size_t)row * im.stride_blocks[c] + col) * 64;
    };
    if(kWrite) {
        const int mcu = (int)(blk / (uint32_t)p.blocks_per_mcu);
        bi = (int)(blk - (uint32_t)mcu * p.blocks_per_mcu);
        mrow = mcu / p.mcus_per_row;
        mcol = mcu - mrow * p.mcus_per_row;
        dst = blk < (uint32_t)p.nblk ? locate() : nullptr;
    }
    if(s.p >= end) return 0;
    BitReader br;
    br.start(words, s.p);
    do {
        const int c = p.bcomp[s.b];
        if(s.z == 0) {
            const int sym = decode_symbol(tab[p.dc_tbl[c]], br, s.bad);
            const int n = sym & 15;
            if(sym > 15) s.bad = 1;
            int diff = 0;
            if(n) {
                diff = extend(br.peek32() >> (32 - n), n);
                br.skip(n);
            }
            if(kWrite && dst) dst[0] = (int16_t)diff;
            s.z = 1;
        }
        else {
            const int sym = decode_symbol(tab[4 + p.ac_tbl[c]], br, s.bad);
            const int r = sym >> 4, n = sym & 15;
            if(n == 0) {
                if(r == 15) {
                    s.z += 16;
                    if(s.z > 63) s.bad = 1, s.z = 64;
                }
                else s.z = 64; // end of block
            }
            else {
                s.z += r;
                if(s.z > 63) s.bad = 1, s.z = 64;
                else {
                    const int val = extend(br.peek32() >> (32 - n), n);
                    br.skip(n);
                    if(kWrite && dst) dst[c_natural[s.z]] = (int16_t)val;
                    s.z++;
                }
            }
        }
        if(s.z >= 64) {
            s.z = 0;
            s.b = s.b + 1 == p.blocks_per_mcu ? 0 : s.b + 1;
            done++;
            if(kWrite) {
                if(blk + done >= (uint32_t)p.nblk) break; // the frame's last block: what follows (padding bits, EOI) is not ours
                if(++bi == p.blocks_per_mcu) {
                    bi = 0;
                    if(++mcol == p.mcus_per_row) mcol = 0, mrow++;
                }
                dst = locate();
            }
        }
    } while(br.pos() < end);
    s.p = br.pos();
    return done;
}

__global__ void __launch_bounds__(kDecThreads) k5_decode_kernel(const DecParams p) {
    __shared__ DecTable           s_tab[8];
    __shared__ unsigned long long s_warp[kDecThreads / 32];
    __shared__ uint32_t           s_total_bits;
    const int                     img = blockIdx.x, tid = threadIdx.x;
    for(int i = tid; i < (int)(sizeof(s_tab) / 4); i += kDecThreads) reinterpret_cast<uint32_t *>(s_tab)[i] = reinterpret_cast<const uint32_t *>(p.tables)[i];
    const mjx_image_desc_t &im = p.items[img];
    const unsigned char    *src = p.data + p.offsets[img];
    const uint32_t          len = p.lengths[img];
    uint32_t               *words = p.words + (size_t)img * p.words_stride;
    unsigned char          *wbytes = reinterpret_cast<unsigned char *>(words);
    unsigned long long      total;

    // ---- 0. the planes start out as zeros (only non-zero coefficients are written) ----
    for(int c = 0; c < p.ncomp; c++) {
        uint4       *pl = reinterpret_cast<uint4 *>(im.plane[c]);
        const size_t n16 = (size_t)im.rows[c] * im.stride_blocks[c] * 8;
        if(pl)
            for(size_t i = tid; i < n16; i += kDecThreads) pl[i] = make_uint4(0u, 0u, 0u, 0u);
    }

    // ---- 1. un-stuff ----
    {
        // every thread a piece of whole 16-byte chunks; a segment that starts on a 16-byte boundary (the batch pipeline's do) is
        // read 128 bits at a time, any other byte by byte
        const uint32_t piece = (((len + kDecThreads - 1) / kDecThreads) + 15u) & ~15u, lo = min(len, (uint32_t)tid * piece), hi = min(len, lo + piece);
        const bool     vec = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
        uint32_t       keep = 0;
        unsigned       prev = lo > 0 ? src[lo - 1] : 0u;
        // kCount: how many bytes stay; otherwise: put them where they belong
        auto pass = [&](const bool count, uint32_t at) -> uint32_t {
            unsigned pv = prev;
            uint32_t i = lo;
            auto     one = [&](unsigned b) {
                const bool drop = b == 0u && pv == 0xffu;
                pv = b;
                if(drop) return;
                if(!count) wbytes[at ^ 3u] = (unsigned char)b; // big-endian inside each 32-bit word
                at++;
            };
            if(vec)
                for(; i + 16 <= hi; i += 16) {
                    const uint4    q = __ldg(reinterpret_cast<const uint4 *>(src + i));
                    const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for(int k = 0; k < 16; k++) one((w4[k >> 2] >> (8 * (k & 3))) & 0xffu);
                }
            for(; i < hi; i++) one(src[i]);
            return at;
        };
        keep = pass(true, 0u);
        const uint32_t at0 = (uint32_t)cta_exclusive_scan(keep, s_warp, &total);
        pass(false, at0);
        // zeros behind the end: a bit window may reach 8 bytes past it
        const uint32_t nbytes = (uint32_t)total;
        if(tid < 16) wbytes[(nbytes + (uint32_t)tid) ^ 3u] = 0;
        if(tid == 0) s_total_bits = nbytes * 8u;
    }
    __syncthreads();
    const uint32_t total_bits = s_total_bits;
    const uint32_t sub_bits = max((uint32_t)kMinSubBits, ((total_bits + kDecThreads - 1) / kDecThreads + 31u) & ~31u);
    const uint32_t nsub = (total_bits + sub_bits - 1) / sub_bits;
    uint2         *entry = p.entry + (size_t)img * p.sub_stride, *exits = p.exits + (size_t)img * p.sub_stride;
    uint32_t      *cnt = p.cnt + (size_t)img * p.sub_stride;
    unsigned char *dirty = p.dirty + (size_t)img * 2 * p.sub_stride, *next = dirty + p.sub_stride;

    // ---- 2. entry states by rounds ----
    for(uint32_t t = tid; t < nsub; t += kDecThreads) {
        DecState s;
        s.p = t * sub_bits, s.b = 0, s.z = 0, s.bad = 0;
        entry[t] = pack_state(s);
        dirty[t] = 1;
        next[t] = 0;
    }
    __syncthreads();
    uint32_t rounds = 0;
    for(;; rounds++) {
        for(uint32_t t = tid; t < nsub; t += kDecThreads) {
            if(!dirty[t]) continue;
            DecState s = unpack_state(entry[t]);
            cnt[t] = decode_run<false>(p, im, s_tab, words, s, min((t + 1) * sub_bits, total_bits), 0u);
            exits[t] = pack_state(s);
        }
        __syncthreads();
        // the thread that owns subsequence t owns entry[t + 1] and next[t + 1] in this phase
        int changed = 0;
        for(uint32_t t = tid; t < nsub; t += kDecThreads) {
            if(!dirty[t] || t + 1 >= nsub) continue;
            const uint2 e = exits[t], old = entry[t + 1];
            if(e.x != old.x || e.y != old.y) {
                entry[t + 1] = e;
                next[t + 1] = 1;
                changed = 1;
            }
        }
        __syncthreads();
        for(uint32_t t = tid; t < nsub; t += kDecThreads) {
            dirty[t] = next[t];
            next[t] = 0;
        }
        if(!__syncthreads_or(changed)) break;
        if(rounds > nsub + 2) break; // (cannot happen: every round fixes at least one more subsequence)
    }
    if(tid == 0) p.rounds[img] = rounds + 1;

    // ---- 3. block index at the start of every subsequence ----
    {
        const uint32_t piece = (nsub + kDecThreads - 1) / kDecThreads, lo = min(nsub, (uint32_t)tid * piece), hi = min(nsub, lo + piece);
        unsigned long long sum = 0;
        for(uint32_t t = lo; t < hi; t++) sum += cnt[t];
        unsigned long long run = cta_exclusive_scan(sum, s_warp, &total);
        for(uint32_t t = lo; t < hi; t++) {
            const uint32_t c = cnt[t];
            cnt[t] = (uint32_t)run;
            run += c;
        }
    }
    __syncthreads();
    int bad = total < (unsigned long long)p.nblk; // fewer blocks in the stream than the frame has

    // ---- 4. decode once more, writing ----
    for(uint32_t t = tid; t < nsub; t += kDecThreads) {
        const uint32_t blk = cnt[t];
        if(blk >= (uint32_t)p.nblk) continue; // behind the last block of the frame: padding bits, EOI, whatever follows
        DecState s = unpack_state(entry[t]);
        const uint32_t done = decode_run<true>(p, im, s_tab, words, s, min((t + 1) * sub_bits, total_bits), blk);
        (void)done;
        if(s.bad) bad = 1;
    }
    bad = __syncthreads_or(bad);
    if(bad) {
        if(tid == 0) p.status[img] = 1u;
        return;
    }

    // ---- 5. DC differences -> DC values, per component in scan order ----
    const int nmcu = p.mcus_per_row * p.mcu_rows;
    for(int c = 0; c < p.ncomp; c++) {
        const int      per = p.h[c] * p.v[c];
        const uint32_t nb = (uint32_t)nmcu * (uint32_t)per;
        const uint32_t piece = (nb + kDecThreads - 1) / kDecThreads, lo = min(nb, (uint32_t)tid * piece), hi = min(nb, lo + piece);
        auto           dc_ptr = [&](uint32_t j) -> int16_t * {
            const int mcu = (int)(j / (uint32_t)per), kk = (int)(j - (uint32_t)mcu * per), mrow = mcu / p.mcus_per_row, mcol = mcu - mrow * p.mcus_per_row;
            const int row = mrow * p.v[c] + kk / p.h[c], col = mcol * p.h[c] + kk % p.h[c];
            if(row >= im.rows[c] || col >= im.stride_blocks[c]) return nullptr;
            return reinterpret_cast<int16_t *>(im.plane[c]) + ((size_t)row * im.stride_blocks[c] + col) * 64;
        };
        long long sum = 0;
        for(uint32_t j = lo; j < hi; j++) {
            const int16_t *q = dc_ptr(j);
            if(q) sum += *q;
        }
        long long run = (long long)cta_exclusive_scan((unsigned long long)sum, s_warp, &total);
        for(uint32_t j = lo; j < hi; j++) {
            int16_t *q = dc_ptr(j);
            if(q) {
                run += *q;
                *q = (int16_t)run;
            }
        }
    }
}

// decoding form of a DHT table; false: not a valid table
static bool derive_dec_table(const mjx_huff_table_t &t, DecTable *d) {
    memset(d, 0, sizeof(*d));
    int  nsym = 0;
    for(int l = 1; l <= 16; l++) nsym += t.bits[l];
    if(nsym > 256) return false;
    memcpy(d->vals, t.vals, 256);
    int code = 0, k = 0;
    for(int l = 1; l <= 16; l++) {
        const int cnt = t.bits[l];
        if(cnt) {
            d->valoff[l] = k - code;
            for(int i = 0; i < cnt; i++, k++, code++) {
                if(code >= (1 << l)) return false; // more codes of this length than there are bit patterns: not a Huffman table
                if(l <= kLook) { // every kLook-bit pattern that starts with this code
                    const int fill = 1 << (kLook - l);
                    for(int f = 0; f < fill; f++) d->look[(code << (kLook - l)) | f] = (uint16_t)((l << 8) | t.vals[k]);
                }
            }
            d->maxcode[l] = code - 1;
            if(code > (1 << l)) return false;
        }
        else d->maxcode[l] = -1;
        code <<= 1;
    }
    d->maxcode[17] = 0x7fffffff;
    return true;
}

} // namespace mjx

using namespace mjx;

extern "C" {

int mjx_huffman_decode_batch_device(mjx_ctx *ctx, const void *data_dev, const uint64_t *offsets, const uint32_t *lengths, int n, const mjx_scan_t *scan,
                                    const mjx_image_desc_t *items_dev, uint32_t *status_dev) {
    if(!ctx) return MJX_ERR_ARG;
    MJX_CUDA(ctx, cudaSetDevice(ctx->device));
    if(!data_dev || !offsets || !lengths || !scan || !items_dev || !status_dev || n < 0) return MJX_ERR_ARG;
    if(n == 0) return MJX_OK;
    if(scan->ncomp < 1 || scan->ncomp > MJX_MAX_COMPONENTS || scan->mcus_per_row < 1 || scan->mcu_rows < 1) return MJX_ERR_ARG;
    DecParams p;
    memset(&p, 0, sizeof(p));
    p.items = items_dev;
    p.data = (const unsigned char *)data_dev;
    p.n = n;
    p.ncomp = scan->ncomp;
    p.mcus_per_row = scan->mcus_per_row;
    p.mcu_rows = scan->mcu_rows;
    int bpm = 0;
    for(int c = 0; c < scan->ncomp; c++) {
        const int h = scan->ncomp == 1 ? 1 : scan->h_samp[c], v = scan->ncomp == 1 ? 1 : scan->v_samp[c];
        if(h < 1 || v < 1 || h > 4 || v > 4 || scan->dc_tbl[c] < 0 || scan->dc_tbl[c] > 3 || scan->ac_tbl[c] < 0 || scan->ac_tbl[c] > 3) return MJX_ERR_ARG;
        p.h[c] = h, p.v[c] = v;
        p.dc_tbl[c] = scan->dc_tbl[c], p.ac_tbl[c] = scan->ac_tbl[c];
        for(int k = 0; k < h * v; k++) {
            if(bpm >= 10) return MJX_ERR_UNSUPPORTED; // D_MAX_BLOCKS_IN_MCU
            p.bcomp[bpm] = (signed char)c, p.bidx[bpm] = (signed char)k;
            p.byoff[bpm] = (signed char)(k / h), p.bxoff[bpm] = (signed char)(k % h);
            bpm++;
        }
    }
    p.blocks_per_mcu = bpm;
    const long long nblk = (long long)scan->mcus_per_row * scan->mcu_rows * bpm;
    if(nblk > 0x3fffffffLL) return MJX_ERR_UNSUPPORTED;
    p.nblk = (int)nblk;
    uint32_t maxlen = 0;
    for(int i = 0; i < n; i++) {
        if(lengths[i] > 0x1ffffff0u) return MJX_ERR_UNSUPPORTED; // the kernel counts bits in 32 bits
        if(lengths[i] > maxlen) maxlen = lengths[i];
    }
    // scratch: tables, offsets, lengths, rounds, then per image the word stream and the subsequence arrays
    const size_t words_stride = ((size_t)maxlen + 3) / 4 + 8, sub_stride = kDecThreads + 8; // at most one subsequence per thread
    size_t       off = 0;
    auto         take = [&](size_t b) {
        const size_t o = off;
        off = (off + b + 255) / 256 * 256;
        return o;
    };
    const size_t o_tab = take(sizeof(DecTable) * 8), o_off = take((size_t)n * 8), o_len = take((size_t)n * 4), o_rounds = take((size_t)n * 4);
    const size_t o_words = take((size_t)n * words_stride * 4), o_entry = take((size_t)n * sub_stride * 8), o_exit = take((size_t)n * sub_stride * 8);
    const size_t o_cnt = take((size_t)n * sub_stride * 4), o_dirty = take((size_t)n * sub_stride * 2);
    if(ctx->huff_bytes < off) {
        MJX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if(ctx->huff) cudaFree(ctx->huff);
        ctx->huff = nullptr;
        ctx->huff_bytes = 0;
        MJX_CUDA(ctx, cudaMalloc(&ctx->huff, off + off / 8 + 4096));
        ctx->huff_bytes = off + off / 8 + 4096;
    }
    char *base = (char *)ctx->huff;
    DecTable tabs[8];
    for(int i = 0; i < 4; i++)
        if(!derive_dec_table(scan->dc[i], &tabs[i]) || !derive_dec_table(scan->ac[i], &tabs[4 + i])) return MJX_ERR_ARG;
    cudaStream_t s = ctx->stream;
    MJX_CUDA(ctx, cudaMemcpyAsync(base + o_tab, tabs, sizeof(tabs), cudaMemcpyHostToDevice, s));
    MJX_CUDA(ctx, cudaMemcpyAsync(base + o_off, offsets, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    MJX_CUDA(ctx, cudaMemcpyAsync(base + o_len, lengths, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    MJX_CUDA(ctx, cudaMemsetAsync(status_dev, 0, (size_t)n * 4, s));
    p.tables = (const DecTable *)(base + o_tab);
    p.offsets = (const uint64_t *)(base + o_off);
    p.lengths = (const uint32_t *)(base + o_len);
    p.rounds = (uint32_t *)(base + o_rounds);
    p.words = (uint32_t *)(base + o_words), p.words_stride = words_stride;
    p.entry = (uint2 *)(base + o_entry), p.exits = (uint2 *)(base + o_exit);
    p.cnt = (uint32_t *)(base + o_cnt), p.dirty = (unsigned char *)(base + o_dirty), p.sub_stride = sub_stride;
    p.status = status_dev;
    k5_decode_kernel<<<n, kDecThreads, 0, s>>>(p);
    ctx->launches += 1;
    const cudaError_t e = cudaGetLastError();
    if(e != cudaSuccess) return fail(ctx, e, "k5_decode_kernel");
    if(getenv("MJX_K5_TRACE") != nullptr) { // diagnostics: how many rounds the entry states needed (synchronises the stream)
        uint32_t *r = (uint32_t *)malloc((size_t)n * 4);
        if(r && cudaMemcpyAsync(r, p.rounds, (size_t)n * 4, cudaMemcpyDeviceToHost, s) == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess) {
            unsigned long long sum = 0;
            uint32_t           mx = 0;
            for(int i = 0; i < n; i++) sum += r[i], mx = r[i] > mx ? r[i] : mx;
            fprintf(stderr, "k5_decode_kernel: %d images, at most %u bytes each: rounds mean %.1f, max %u\n", n, maxlen, (double)sum / n, mx);
        }
        free(r);
    }
    return MJX_OK;
}

} // extern "C"
