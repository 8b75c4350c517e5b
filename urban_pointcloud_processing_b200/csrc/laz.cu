// "Next" row 1 -- LAZ ingest and write-back: decoder AND encoder for LASzip-compressed point data,
// host code (the encoder is the second half of this file).
//
// The reference reads its tiles with laspy[lazrs] (las_utils.read_las, reference utils/las_utils.py:118-120;
// every CycloMedia tile and the demo AHN clouds are .laz).  Neither laspy nor a LASzip library is in this
// image, and the format is a published one (M. Isenburg, "LASzip: lossless compression of LiDAR data", PE&RS
// 2013; LASzip 2.x / 3.x open-source reference implementation): an adaptive arithmetic coder (A. Said's FastAC)
// driven by per-field context models.  This file restates the DECODER for the "pointwise chunked" compressor
// (type 2) with the version-2 item codecs of point formats 0-3: POINT10 (20 bytes), GPSTIME11, RGB12 and BYTE
// (extra bytes).  Output = the uncompressed point records, byte for byte what an uncompressed .las holds.
// LAS 1.4 "layered" items (formats 6-10, compressor 3) are not covered.
//
// Conformance is pinned on the two demo clouds of the reference (datasets/ahn/ahn_*.laz, format 1, 43 536 and
// 45 345 points): decoded coordinate ranges equal the header's min / max exactly, the returns histogram equals
// the header's, and the rasters computed from the decoded points equal the ahn_*.npz files the reference
// shipped beside them (tests/test_abi_and_host.py, tests/test_gpu_parity.py).
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <new>
#include <thread>
#include <vector>
#include "common.cuh"

namespace upcp {
namespace laz {

typedef uint8_t U8; typedef uint16_t U16; typedef uint32_t U32; typedef int32_t I32; typedef int64_t I64; typedef uint64_t U64;

constexpr U32 AC_MinLength = 0x01000000u, AC_MaxLength = 0xFFFFFFFFu;
constexpr U32 BM_LengthShift = 13, BM_MaxCount = 1u << BM_LengthShift;
constexpr U32 DM_LengthShift = 15, DM_MaxCount = 1u << DM_LengthShift;

struct ByteStream {
    const U8* p; const U8* end; bool overrun;
    U8 get() { if (p < end) return *p++; overrun = true; return 0; }
};

struct BitModel {
    U32 update_cycle, bits_until_update, bit_0_prob, bit_0_count, bit_count;
    void init() { bit_0_count = 1; bit_count = 2; bit_0_prob = 1u << (BM_LengthShift - 1); update_cycle = bits_until_update = 4; }
    void update() {
        if ((bit_count += update_cycle) > BM_MaxCount) {
            bit_count = (bit_count + 1) >> 1;
            bit_0_count = (bit_0_count + 1) >> 1;
            if (bit_0_count == bit_count) ++bit_count;
        }
        const U32 scale = 0x80000000u / bit_count;
        bit_0_prob = (bit_0_count * scale) >> (31 - BM_LengthShift);
        update_cycle = (5 * update_cycle) >> 2;
        if (update_cycle > 64) update_cycle = 64;
        bits_until_update = update_cycle;
    }
};

struct SymbolModel {
    U32 symbols = 0, last_symbol = 0, table_size = 0, table_shift = 0, total_count = 0, update_cycle = 0, symbols_until_update = 0;
    std::vector<U32> distribution, symbol_count, decoder_table;
    void create(U32 n) { symbols = n; }
    void init() {
        if (distribution.empty()) {
            last_symbol = symbols - 1;
            if (symbols > 16) {
                U32 table_bits = 3;
                while (symbols > (1u << (table_bits + 2))) ++table_bits;
                table_size = 1u << table_bits;
                table_shift = DM_LengthShift - table_bits;
                decoder_table.assign(table_size + 2, 0);
            } else { table_size = table_shift = 0; }
            distribution.assign(symbols, 0);
            symbol_count.assign(symbols, 0);
        }
        total_count = 0;
        update_cycle = symbols;
        for (U32 k = 0; k < symbols; ++k) symbol_count[k] = 1;
        update();
        symbols_until_update = update_cycle = (symbols + 6) >> 1;
    }
    void update() {
        if ((total_count += update_cycle) > DM_MaxCount) {
            total_count = 0;
            for (U32 n = 0; n < symbols; ++n) total_count += (symbol_count[n] = (symbol_count[n] + 1) >> 1);
        }
        U32 sum = 0, s = 0;
        const U32 scale = 0x80000000u / total_count;
        if (table_size == 0) {
            for (U32 k = 0; k < symbols; ++k) { distribution[k] = (scale * sum) >> (31 - DM_LengthShift); sum += symbol_count[k]; }
        } else {
            for (U32 k = 0; k < symbols; ++k) {
                distribution[k] = (scale * sum) >> (31 - DM_LengthShift);
                sum += symbol_count[k];
                const U32 w = distribution[k] >> table_shift;
                while (s < w) decoder_table[++s] = k - 1;
            }
            decoder_table[0] = 0;
            while (s <= table_size) decoder_table[++s] = symbols - 1;
        }
        update_cycle = (5 * update_cycle) >> 2;
        const U32 max_cycle = (symbols + 6) << 3;
        if (update_cycle > max_cycle) update_cycle = max_cycle;
        symbols_until_update = update_cycle;
    }
};

struct Decoder {
    ByteStream* in = nullptr;
    U32 value = 0, length = 0;
    void init(ByteStream* s) {
        in = s;
        length = AC_MaxLength;
        value = ((U32)in->get() << 24); value |= ((U32)in->get() << 16); value |= ((U32)in->get() << 8); value |= (U32)in->get();
    }
    void renorm() { do { value = (value << 8) | in->get(); } while ((length <<= 8) < AC_MinLength); }
    U32 decodeBit(BitModel* m) {
        const U32 x = m->bit_0_prob * (length >> BM_LengthShift);
        const U32 sym = (value >= x);
        if (sym == 0) { length = x; ++m->bit_0_count; } else { value -= x; length -= x; }
        if (length < AC_MinLength) renorm();
        if (--m->bits_until_update == 0) m->update();
        return sym;
    }
    U32 decodeSymbol(SymbolModel* m) {
        U32 n, sym, x, y = length;
        if (m->table_size) {
            const U32 dv = value / (length >>= DM_LengthShift);
            const U32 t = dv >> m->table_shift;
            sym = m->decoder_table[t];
            n = m->decoder_table[t + 1] + 1;
            while (n > sym + 1) {
                const U32 k = (sym + n) >> 1;
                if (m->distribution[k] > dv) n = k; else sym = k;
            }
            x = m->distribution[sym] * length;
            if (sym != m->last_symbol) y = m->distribution[sym + 1] * length;
        } else {
            x = sym = 0;
            length >>= DM_LengthShift;
            U32 k = (n = m->symbols) >> 1;
            do {
                const U32 z = length * m->distribution[k];
                if (z > value) { n = k; y = z; } else { sym = k; x = z; }
            } while ((k = (sym + n) >> 1) != sym);
        }
        value -= x;
        length = y - x;
        if (length < AC_MinLength) renorm();
        ++m->symbol_count[sym];
        if (--m->symbols_until_update == 0) m->update();
        return sym;
    }
    U32 readBits(U32 bits) {
        if (bits > 19) {
            const U32 lo = readShort();
            bits -= 16;
            const U32 hi = readBits(bits) << 16;
            return hi | lo;
        }
        const U32 sym = value / (length >>= bits);
        value -= length * sym;
        if (length < AC_MinLength) renorm();
        return sym;
    }
    U32 readShort() {
        const U32 sym = value / (length >>= 16);
        value -= length * sym;
        if (length < AC_MinLength) renorm();
        return sym & 0xffffu;
    }
    U32 readInt() { const U32 lo = readShort(); const U32 hi = readShort(); return (hi << 16) | lo; }
};

struct IntegerCompressor {
    Decoder* dec = nullptr;
    U32 k = 0, contexts = 1, bits_high = 8, bits = 16, corr_bits = 0, corr_range = 0;
    I32 corr_min = 0, corr_max = 0;
    std::vector<SymbolModel> mBits, mCorrector;      // mCorrector[0] is unused (bit model below)
    BitModel mCorrector0;
    void create(Decoder* d, U32 bits_, U32 contexts_ = 1, U32 bits_high_ = 8) {
        dec = d; bits = bits_; contexts = contexts_; bits_high = bits_high_;
        if (bits && bits < 32) {
            corr_bits = bits; corr_range = 1u << bits;
            corr_min = -((I32)(corr_range / 2)); corr_max = corr_min + (I32)corr_range - 1;
        } else { corr_bits = 32; corr_range = 0; corr_min = INT32_MIN; corr_max = INT32_MAX; }
        mBits.assign(contexts, SymbolModel());
        for (auto& m : mBits) m.create(corr_bits + 1);
        mCorrector.assign(corr_bits + 1, SymbolModel());
        for (U32 i = 1; i <= corr_bits; ++i) mCorrector[i].create(i <= bits_high ? (1u << i) : (1u << bits_high));
    }
    void init() {
        for (auto& m : mBits) m.init();
        mCorrector0.init();
        for (U32 i = 1; i <= corr_bits; ++i) mCorrector[i].init();
    }
    I32 readCorrector(SymbolModel* mb) {
        I32 c;
        k = dec->decodeSymbol(mb);
        if (k) {
            if (k < 32) {
                if (k <= bits_high) c = (I32)dec->decodeSymbol(&mCorrector[k]);
                else {
                    const U32 k1 = k - bits_high;
                    c = (I32)dec->decodeSymbol(&mCorrector[k]);
                    const I32 c1 = (I32)dec->readBits(k1);
                    c = (I32)(((U32)c << k1) | (U32)c1);
                }
                if (c >= (I32)(1u << (k - 1))) c += 1; else c -= (I32)((1u << k) - 1);
            } else c = corr_min;
        } else c = (I32)dec->decodeBit(&mCorrector0);
        return c;
    }
    I32 decompress(I32 pred, U32 context = 0) {
        I32 real = (I32)((U32)pred + (U32)readCorrector(&mBits[context]));
        if (real < 0) real = (I32)((U32)real + corr_range);
        else if ((U32)real >= corr_range) real = (I32)((U32)real - corr_range);
        return real;
    }
    U32 getK() const { return k; }
};

struct StreamingMedian5 {
    I32 values[5]; bool high;
    void init() { values[0] = values[1] = values[2] = values[3] = values[4] = 0; high = true; }
    void add(I32 v) {
        if (high) {
            if (v < values[2]) {
                values[4] = values[3]; values[3] = values[2];
                if (v < values[0]) { values[2] = values[1]; values[1] = values[0]; values[0] = v; }
                else if (v < values[1]) { values[2] = values[1]; values[1] = v; }
                else values[2] = v;
            } else {
                if (v < values[3]) { values[4] = values[3]; values[3] = v; } else values[4] = v;
                high = false;
            }
        } else {
            if (values[2] < v) {
                values[0] = values[1]; values[1] = values[2];
                if (values[4] < v) { values[2] = values[3]; values[3] = values[4]; values[4] = v; }
                else if (values[3] < v) { values[2] = values[3]; values[3] = v; }
                else values[2] = v;
            } else {
                if (values[1] < v) { values[0] = values[1]; values[1] = v; } else values[0] = v;
                high = true;
            }
        }
    }
    I32 get() const { return values[2]; }
};

static const U8 number_return_map[8][8] = {
    {15, 14, 13, 12, 11, 10, 9, 8}, {14, 0, 1, 3, 6, 10, 10, 9}, {13, 1, 2, 4, 7, 11, 11, 10}, {12, 3, 4, 5, 8, 12, 12, 11},
    {11, 6, 7, 8, 9, 13, 13, 12}, {10, 10, 11, 12, 13, 14, 14, 13}, {9, 10, 11, 12, 13, 14, 15, 14}, {8, 9, 10, 11, 12, 13, 14, 15}};
static const U8 number_return_level[8][8] = {
    {0, 1, 2, 3, 4, 5, 6, 7}, {1, 0, 1, 2, 3, 4, 5, 6}, {2, 1, 0, 1, 2, 3, 4, 5}, {3, 2, 1, 0, 1, 2, 3, 4},
    {4, 3, 2, 1, 0, 1, 2, 3}, {5, 4, 3, 2, 1, 0, 1, 2}, {6, 5, 4, 3, 2, 1, 0, 1}, {7, 6, 5, 4, 3, 2, 1, 0}};

static inline U8 u8_fold(I32 n) { return (U8)(n & 0xff); }
static inline U8 u8_clamp(I32 n) { return n <= 0 ? 0 : (n >= 255 ? 255 : (U8)n); }

struct ItemReader {
    virtual ~ItemReader() {}
    virtual void init(const U8* item) = 0;
    virtual void read(U8* item) = 0;
};

struct Point10Reader : ItemReader {
    Decoder* dec;
    U8 last_item[20];
    U16 last_intensity[16];
    StreamingMedian5 last_x_diff_median5[16], last_y_diff_median5[16];
    I32 last_height[8];
    SymbolModel m_changed_values, m_scan_angle_rank[2];
    std::vector<SymbolModel> m_bit_byte, m_classification, m_user_data;
    std::vector<char> has_bit_byte, has_classification, has_user_data;
    IntegerCompressor ic_intensity, ic_point_source_ID, ic_dx, ic_dy, ic_z;
    explicit Point10Reader(Decoder* d) : dec(d) {
        m_changed_values.create(64);
        m_scan_angle_rank[0].create(256); m_scan_angle_rank[1].create(256);
        m_bit_byte.assign(256, SymbolModel()); m_classification.assign(256, SymbolModel()); m_user_data.assign(256, SymbolModel());
        ic_intensity.create(d, 16, 4); ic_point_source_ID.create(d, 16); ic_dx.create(d, 32, 2); ic_dy.create(d, 32, 22);
        ic_z.create(d, 32, 20);
    }
    void init(const U8* item) override {
        for (int i = 0; i < 16; ++i) {
            last_x_diff_median5[i].init(); last_y_diff_median5[i].init();
            last_intensity[i] = 0; last_height[i / 2] = 0;
        }
        m_changed_values.init(); m_scan_angle_rank[0].init(); m_scan_angle_rank[1].init();
        has_bit_byte.assign(256, 0); has_classification.assign(256, 0); has_user_data.assign(256, 0);
        ic_intensity.init(); ic_point_source_ID.init(); ic_dx.init(); ic_dy.init(); ic_z.init();
        memcpy(last_item, item, 20);
        last_item[12] = 0; last_item[13] = 0;         // the intensity of the first point is not a prediction base
    }
    static SymbolModel* lazy(std::vector<SymbolModel>& v, std::vector<char>& has, U32 i) {
        if (!has[i]) { v[i] = SymbolModel(); v[i].create(256); v[i].init(); has[i] = 1; }
        return &v[i];
    }
    void read(U8* item) override {
        U32 r, n, m, l;
        I32 x, y, z;
        U16 intensity, psid;
        memcpy(&x, last_item, 4); memcpy(&y, last_item + 4, 4); memcpy(&z, last_item + 8, 4);
        memcpy(&intensity, last_item + 12, 2); memcpy(&psid, last_item + 18, 2);
        const U32 changed = dec->decodeSymbol(&m_changed_values);
        if (changed) {
            if (changed & 32) last_item[14] = (U8)dec->decodeSymbol(lazy(m_bit_byte, has_bit_byte, last_item[14]));
            r = last_item[14] & 7; n = (last_item[14] >> 3) & 7;
            m = number_return_map[n][r]; l = number_return_level[n][r];
            if (changed & 16) {
                intensity = (U16)ic_intensity.decompress(last_intensity[m], (m < 3 ? m : 3));
                last_intensity[m] = intensity;
            } else intensity = last_intensity[m];
            if (changed & 8) last_item[15] = (U8)dec->decodeSymbol(lazy(m_classification, has_classification, last_item[15]));
            if (changed & 4) {
                const I32 val = (I32)dec->decodeSymbol(&m_scan_angle_rank[(last_item[14] >> 6) & 1]);
                last_item[16] = u8_fold(val + last_item[16]);
            }
            if (changed & 2) last_item[17] = (U8)dec->decodeSymbol(lazy(m_user_data, has_user_data, last_item[17]));
            if (changed & 1) psid = (U16)ic_point_source_ID.decompress(psid);
        } else {
            r = last_item[14] & 7; n = (last_item[14] >> 3) & 7;
            m = number_return_map[n][r]; l = number_return_level[n][r];
        }
        I32 median = last_x_diff_median5[m].get();
        I32 diff = ic_dx.decompress(median, n == 1);
        x = (I32)((U32)x + (U32)diff);
        last_x_diff_median5[m].add(diff);
        median = last_y_diff_median5[m].get();
        U32 k_bits = ic_dx.getK();
        diff = ic_dy.decompress(median, (n == 1) + (k_bits < 20 ? (k_bits & ~1u) : 20));
        y = (I32)((U32)y + (U32)diff);
        last_y_diff_median5[m].add(diff);
        k_bits = (ic_dx.getK() + ic_dy.getK()) / 2;
        z = ic_z.decompress(last_height[l], (n == 1) + (k_bits < 18 ? (k_bits & ~1u) : 18));
        last_height[l] = z;
        memcpy(last_item, &x, 4); memcpy(last_item + 4, &y, 4); memcpy(last_item + 8, &z, 4);
        memcpy(last_item + 12, &intensity, 2); memcpy(last_item + 18, &psid, 2);
        memcpy(item, last_item, 20);
    }
};

struct GpsTime11Reader : ItemReader {
    static constexpr I32 MULTI = 500, MULTI_MINUS = -10, MULTI_UNCHANGED = MULTI - MULTI_MINUS + 1,
                         MULTI_CODE_FULL = MULTI - MULTI_MINUS + 2, MULTI_TOTAL = MULTI - MULTI_MINUS + 6;
    Decoder* dec;
    U32 last = 0, next = 0;
    U64 last_gpstime[4];
    I32 last_gpstime_diff[4], multi_extreme_counter[4];
    SymbolModel m_multi, m_0diff;
    IntegerCompressor ic;
    explicit GpsTime11Reader(Decoder* d) : dec(d) { m_multi.create(MULTI_TOTAL); m_0diff.create(6); ic.create(d, 32, 9); }
    void init(const U8* item) override {
        last = next = 0;
        for (int i = 0; i < 4; ++i) { last_gpstime_diff[i] = 0; multi_extreme_counter[i] = 0; last_gpstime[i] = 0; }
        m_multi.init(); m_0diff.init(); ic.init();
        memcpy(&last_gpstime[0], item, 8);
    }
    void full() {
        next = (next + 1) & 3;
        last_gpstime[next] = (U64)(U32)ic.decompress((I32)(last_gpstime[last] >> 32), 8);
        last_gpstime[next] = last_gpstime[next] << 32;
        last_gpstime[next] |= (U64)dec->readInt();
        last = next;
        last_gpstime_diff[last] = 0;
        multi_extreme_counter[last] = 0;
    }
    void read(U8* item) override {
        I32 multi;
        if (last_gpstime_diff[last] == 0) {
            multi = (I32)dec->decodeSymbol(&m_0diff);
            if (multi == 1) {
                last_gpstime_diff[last] = ic.decompress(0, 0);
                last_gpstime[last] = (U64)((I64)last_gpstime[last] + last_gpstime_diff[last]);
                multi_extreme_counter[last] = 0;
            } else if (multi == 2) {
                full();
            } else if (multi > 2) {
                last = (last + (U32)multi - 2) & 3;
                read(item);
                return;
            }
        } else {
            multi = (I32)dec->decodeSymbol(&m_multi);
            if (multi == 1) {
                last_gpstime[last] = (U64)((I64)last_gpstime[last] + ic.decompress(last_gpstime_diff[last], 1));
                multi_extreme_counter[last] = 0;
            } else if (multi < MULTI_UNCHANGED) {
                I32 gpstime_diff;
                if (multi == 0) {
                    gpstime_diff = ic.decompress(0, 7);
                    if (++multi_extreme_counter[last] > 3) { last_gpstime_diff[last] = gpstime_diff; multi_extreme_counter[last] = 0; }
                } else if (multi < MULTI) {
                    gpstime_diff = ic.decompress((I32)((U32)multi * (U32)last_gpstime_diff[last]), multi < 10 ? 2 : 3);
                } else if (multi == MULTI) {
                    gpstime_diff = ic.decompress((I32)((U32)MULTI * (U32)last_gpstime_diff[last]), 4);
                    if (++multi_extreme_counter[last] > 3) { last_gpstime_diff[last] = gpstime_diff; multi_extreme_counter[last] = 0; }
                } else {
                    multi = MULTI - multi;
                    if (multi > MULTI_MINUS) {
                        gpstime_diff = ic.decompress((I32)((U32)multi * (U32)last_gpstime_diff[last]), 5);
                    } else {
                        gpstime_diff = ic.decompress((I32)((U32)MULTI_MINUS * (U32)last_gpstime_diff[last]), 6);
                        if (++multi_extreme_counter[last] > 3) { last_gpstime_diff[last] = gpstime_diff; multi_extreme_counter[last] = 0; }
                    }
                }
                last_gpstime[last] = (U64)((I64)last_gpstime[last] + gpstime_diff);
            } else if (multi == MULTI_CODE_FULL) {
                full();
            } else if (multi >= MULTI_CODE_FULL) {
                last = (last + (U32)multi - (U32)MULTI_CODE_FULL) & 3;
                read(item);
                return;
            }
        }
        memcpy(item, &last_gpstime[last], 8);
    }
};

struct Rgb12Reader : ItemReader {
    Decoder* dec;
    U16 last_item[3];
    SymbolModel m_byte_used, m_diff[6];
    explicit Rgb12Reader(Decoder* d) : dec(d) { m_byte_used.create(128); for (auto& m : m_diff) m.create(256); }
    void init(const U8* item) override { m_byte_used.init(); for (auto& m : m_diff) m.init(); memcpy(last_item, item, 6); }
    void read(U8* out) override {
        U16 item[3];
        U8 corr;
        I32 diff = 0;
        const U32 sym = dec->decodeSymbol(&m_byte_used);
        if (sym & 1) { corr = (U8)dec->decodeSymbol(&m_diff[0]); item[0] = (U16)u8_fold(corr + (last_item[0] & 255)); }
        else item[0] = last_item[0] & 0xFF;
        if (sym & 2) { corr = (U8)dec->decodeSymbol(&m_diff[1]); item[0] |= ((U16)u8_fold(corr + (last_item[0] >> 8))) << 8; }
        else item[0] |= (last_item[0] & 0xFF00);
        if (sym & 64) {
            diff = (item[0] & 0x00FF) - (last_item[0] & 0x00FF);
            if (sym & 4) { corr = (U8)dec->decodeSymbol(&m_diff[2]); item[1] = (U16)u8_fold(corr + u8_clamp(diff + (last_item[1] & 255))); }
            else item[1] = last_item[1] & 0xFF;
            if (sym & 16) {
                corr = (U8)dec->decodeSymbol(&m_diff[4]);
                diff = (diff + ((item[1] & 0x00FF) - (last_item[1] & 0x00FF))) / 2;
                item[2] = (U16)u8_fold(corr + u8_clamp(diff + (last_item[2] & 255)));
            } else item[2] = last_item[2] & 0xFF;
            diff = (item[0] >> 8) - (last_item[0] >> 8);
            if (sym & 8) { corr = (U8)dec->decodeSymbol(&m_diff[3]); item[1] |= ((U16)u8_fold(corr + u8_clamp(diff + (last_item[1] >> 8)))) << 8; }
            else item[1] |= (last_item[1] & 0xFF00);
            if (sym & 32) {
                corr = (U8)dec->decodeSymbol(&m_diff[5]);
                diff = (diff + ((item[1] >> 8) - (last_item[1] >> 8))) / 2;
                item[2] |= ((U16)u8_fold(corr + u8_clamp(diff + (last_item[2] >> 8)))) << 8;
            } else item[2] |= (last_item[2] & 0xFF00);
        } else { item[1] = item[0]; item[2] = item[0]; }
        memcpy(last_item, item, 6);
        memcpy(out, item, 6);
    }
};

struct ByteReader : ItemReader {
    Decoder* dec;
    U32 number;
    std::vector<U8> last_item;
    std::vector<SymbolModel> m_byte;
    ByteReader(Decoder* d, U32 n) : dec(d), number(n), last_item(n), m_byte(n) { for (auto& m : m_byte) m.create(256); }
    void init(const U8* item) override { for (auto& m : m_byte) m.init(); memcpy(last_item.data(), item, number); }
    void read(U8* item) override {
        for (U32 i = 0; i < number; ++i) { item[i] = u8_fold(last_item[i] + (I32)dec->decodeSymbol(&m_byte[i])); }
        memcpy(last_item.data(), item, number);
    }
};


// ------------------------------------------------------------------------------------------------ encoder ----
// The inverse of everything above, for las_utils.label_and_save_las_as_laz (reference utils/las_utils.py:133-164:
// laspy writes `.laz` through lazrs).  Same models, same contexts, same state updates as the readers -- the
// arithmetic ENCODER of FastAC (base / length interval, carry propagation into the bytes already written,
// byte-wise renormalisation, and a final flush that leaves the stream exactly as long as the decoder reads).
// Pinned on the reference's own data: re-encoding the decoded demo clouds reproduces their compressed point
// data byte for byte (tests/test_abi_and_host.py).
struct Encoder {
    std::vector<U8>* out = nullptr;
    U32 base = 0, length = 0;
    void init(std::vector<U8>* o) { out = o; base = 0; length = AC_MaxLength; }
    void propagate_carry() {
        size_t i = out->size();
        while (i > 0 && (*out)[i - 1] == 0xFFu) { (*out)[i - 1] = 0; --i; }
        if (i > 0) ++(*out)[i - 1];
    }
    void renorm() { do { out->push_back((U8)(base >> 24)); base <<= 8; } while ((length <<= 8) < AC_MinLength); }
    void encodeBit(BitModel* m, U32 sym) {
        const U32 x = m->bit_0_prob * (length >> BM_LengthShift);
        if (sym == 0) { length = x; ++m->bit_0_count; }
        else { const U32 init_base = base; base += x; length -= x; if (init_base > base) propagate_carry(); }
        if (length < AC_MinLength) renorm();
        if (--m->bits_until_update == 0) m->update();
    }
    void encodeSymbol(SymbolModel* m, U32 sym) {
        U32 x;
        const U32 init_base = base;
        if (sym == m->last_symbol) {
            x = m->distribution[sym] * (length >> DM_LengthShift);
            base += x; length -= x;
        } else {
            x = m->distribution[sym] * (length >>= DM_LengthShift);
            base += x; length = m->distribution[sym + 1] * length - x;
        }
        if (init_base > base) propagate_carry();
        if (length < AC_MinLength) renorm();
        ++m->symbol_count[sym];
        if (--m->symbols_until_update == 0) m->update();
    }
    void writeShort(U32 sym) {
        const U32 init_base = base;
        base += (sym & 0xffffu) * (length >>= 16);
        if (init_base > base) propagate_carry();
        if (length < AC_MinLength) renorm();
    }
    void writeBits(U32 bits, U32 sym) {
        if (bits > 19) { writeShort(sym & 0xffffu); sym >>= 16; bits -= 16; }
        const U32 init_base = base;
        base += sym * (length >>= bits);
        if (init_base > base) propagate_carry();
        if (length < AC_MinLength) renorm();
    }
    void writeInt(U32 sym) { writeShort(sym & 0xffffu); writeShort(sym >> 16); }
    void done() {
        const U32 init_base = base;
        bool another_byte = true;
        if (length > 2 * AC_MinLength) { base += AC_MinLength; length = AC_MinLength >> 1; }
        else { base += AC_MinLength >> 1; length = AC_MinLength >> 9; another_byte = false; }
        if (init_base > base) propagate_carry();
        renorm();
        // the decoder reads four bytes ahead: pad so that it ends exactly where the stream does
        out->push_back(0); out->push_back(0);
        if (another_byte) out->push_back(0);
    }
};

// IntegerCompressor::compress (shares the models of the struct above through a decoder-less instance)
struct IntegerEncoder {
    Encoder* enc = nullptr;
    IntegerCompressor m;             // models, ranges and the last k
    void create(Encoder* e, U32 bits, U32 contexts = 1, U32 bits_high = 8) { enc = e; m.create(nullptr, bits, contexts, bits_high); }
    void init() { m.init(); }
    U32 getK() const { return m.k; }
    void compress(I32 pred, I32 real, U32 context = 0) {
        I32 corr = (I32)((U32)real - (U32)pred);
        if (m.corr_range) {
            if (corr < m.corr_min) corr = (I32)((U32)corr + m.corr_range);
            else if (corr > m.corr_max) corr = (I32)((U32)corr - m.corr_range);
        }
        SymbolModel* mb = &m.mBits[context];
        // the tightest interval [-(2^k - 1), 2^k] that holds the corrector
        U32 c1 = corr <= 0 ? (U32)0 - (U32)corr : (U32)corr - 1u;
        U32 k = 0;
        while (c1) { c1 >>= 1; ++k; }
        m.k = k;
        enc->encodeSymbol(mb, k);
        if (k) {
            if (k < 32) {
                U32 c = corr < 0 ? (U32)corr + ((1u << k) - 1u) : (U32)corr - 1u;      // into [0, 2^k - 1]
                if (k <= m.bits_high) enc->encodeSymbol(&m.mCorrector[k], c);
                else {
                    const U32 k1 = k - m.bits_high;
                    const U32 lo = c & ((1u << k1) - 1u);
                    c >>= k1;
                    enc->encodeSymbol(&m.mCorrector[k], c);
                    enc->writeBits(k1, lo);
                }
            }
        } else enc->encodeBit(&m.mCorrector0, (U32)corr);
    }
};

struct ItemWriter {
    virtual ~ItemWriter() {}
    virtual void init(const U8* item) = 0;
    virtual void write(const U8* item) = 0;
};

struct Point10Writer : ItemWriter {
    Encoder* enc;
    U8 last_item[20];
    U16 last_intensity[16];
    StreamingMedian5 last_x_diff_median5[16], last_y_diff_median5[16];
    I32 last_height[8];
    SymbolModel m_changed_values, m_scan_angle_rank[2];
    std::vector<SymbolModel> m_bit_byte, m_classification, m_user_data;
    std::vector<char> has_bit_byte, has_classification, has_user_data;
    IntegerEncoder ic_intensity, ic_point_source_ID, ic_dx, ic_dy, ic_z;
    explicit Point10Writer(Encoder* e) : enc(e) {
        m_changed_values.create(64);
        m_scan_angle_rank[0].create(256); m_scan_angle_rank[1].create(256);
        m_bit_byte.assign(256, SymbolModel()); m_classification.assign(256, SymbolModel()); m_user_data.assign(256, SymbolModel());
        ic_intensity.create(e, 16, 4); ic_point_source_ID.create(e, 16); ic_dx.create(e, 32, 2); ic_dy.create(e, 32, 22);
        ic_z.create(e, 32, 20);
    }
    void init(const U8* item) override {
        for (int i = 0; i < 16; ++i) {
            last_x_diff_median5[i].init(); last_y_diff_median5[i].init();
            last_intensity[i] = 0; last_height[i / 2] = 0;
        }
        m_changed_values.init(); m_scan_angle_rank[0].init(); m_scan_angle_rank[1].init();
        has_bit_byte.assign(256, 0); has_classification.assign(256, 0); has_user_data.assign(256, 0);
        ic_intensity.init(); ic_point_source_ID.init(); ic_dx.init(); ic_dy.init(); ic_z.init();
        memcpy(last_item, item, 20);
        last_item[12] = 0; last_item[13] = 0;
    }
    void write(const U8* item) override {
        I32 x, y, z, lx, ly;
        U16 intensity, psid, lpsid;
        memcpy(&x, item, 4); memcpy(&y, item + 4, 4); memcpy(&z, item + 8, 4);
        memcpy(&intensity, item + 12, 2); memcpy(&psid, item + 18, 2);
        memcpy(&lx, last_item, 4); memcpy(&ly, last_item + 4, 4); memcpy(&lpsid, last_item + 18, 2);
        const U32 r = item[14] & 7, n = (item[14] >> 3) & 7;
        const U32 m = number_return_map[n][r], l = number_return_level[n][r];
        const U32 changed = ((U32)(last_item[14] != item[14]) << 5) | ((U32)(last_intensity[m] != intensity) << 4) |
                            ((U32)(last_item[15] != item[15]) << 3) | ((U32)(last_item[16] != item[16]) << 2) |
                            ((U32)(last_item[17] != item[17]) << 1) | (U32)(lpsid != psid);
        enc->encodeSymbol(&m_changed_values, changed);
        if (changed & 32) enc->encodeSymbol(Point10Reader::lazy(m_bit_byte, has_bit_byte, last_item[14]), item[14]);
        if (changed & 16) { ic_intensity.compress(last_intensity[m], intensity, (m < 3 ? m : 3)); last_intensity[m] = intensity; }
        if (changed & 8) enc->encodeSymbol(Point10Reader::lazy(m_classification, has_classification, last_item[15]), item[15]);
        if (changed & 4) enc->encodeSymbol(&m_scan_angle_rank[(item[14] >> 6) & 1], u8_fold((I32)item[16] - (I32)last_item[16]));
        if (changed & 2) enc->encodeSymbol(Point10Reader::lazy(m_user_data, has_user_data, last_item[17]), item[17]);
        if (changed & 1) ic_point_source_ID.compress(lpsid, psid);
        I32 median = last_x_diff_median5[m].get();
        I32 diff = (I32)((U32)x - (U32)lx);
        ic_dx.compress(median, diff, n == 1);
        last_x_diff_median5[m].add(diff);
        U32 k_bits = ic_dx.getK();
        median = last_y_diff_median5[m].get();
        diff = (I32)((U32)y - (U32)ly);
        ic_dy.compress(median, diff, (n == 1) + (k_bits < 20 ? (k_bits & ~1u) : 20));
        last_y_diff_median5[m].add(diff);
        k_bits = (ic_dx.getK() + ic_dy.getK()) / 2;
        ic_z.compress(last_height[l], z, (n == 1) + (k_bits < 18 ? (k_bits & ~1u) : 18));
        last_height[l] = z;
        memcpy(last_item, item, 20);
    }
};

struct GpsTime11Writer : ItemWriter {
    typedef GpsTime11Reader G;
    Encoder* enc;
    U32 last = 0, next = 0;
    U64 last_gpstime[4];
    I32 last_gpstime_diff[4], multi_extreme_counter[4];
    SymbolModel m_multi, m_0diff;
    IntegerEncoder ic;
    explicit GpsTime11Writer(Encoder* e) : enc(e) { m_multi.create(G::MULTI_TOTAL); m_0diff.create(6); ic.create(e, 32, 9); }
    void init(const U8* item) override {
        last = next = 0;
        for (int i = 0; i < 4; ++i) { last_gpstime_diff[i] = 0; multi_extreme_counter[i] = 0; last_gpstime[i] = 0; }
        m_multi.init(); m_0diff.init(); ic.init();
        memcpy(&last_gpstime[0], item, 8);
    }
    static bool fits32(I64 d) { return d == (I64)(I32)d; }
    void bump(I32 diff) { if (++multi_extreme_counter[last] > 3) { last_gpstime_diff[last] = diff; multi_extreme_counter[last] = 0; } }
    // the value belongs to another of the four running sequences, or starts a new one
    void far_away(U64 t, SymbolModel* model, U32 code_full) {
        for (U32 i = 1; i < 4; ++i) {
            if (fits32((I64)(t - last_gpstime[(last + i) & 3]))) {
                enc->encodeSymbol(model, code_full + i);
                last = (last + i) & 3;
                write_time(t);
                return;
            }
        }
        enc->encodeSymbol(model, code_full);
        ic.compress((I32)(last_gpstime[last] >> 32), (I32)(t >> 32), 8);
        enc->writeInt((U32)t);
        next = (next + 1) & 3;
        last = next;
        last_gpstime_diff[last] = 0;
        multi_extreme_counter[last] = 0;
        last_gpstime[last] = t;
    }
    void write_time(U64 t) {
        if (last_gpstime_diff[last] == 0) {
            if (t == last_gpstime[last]) { enc->encodeSymbol(&m_0diff, 0); return; }
            const I64 d64 = (I64)(t - last_gpstime[last]);
            if (fits32(d64)) {
                enc->encodeSymbol(&m_0diff, 1);
                ic.compress(0, (I32)d64, 0);
                last_gpstime_diff[last] = (I32)d64;
                multi_extreme_counter[last] = 0;
                last_gpstime[last] = t;
            } else far_away(t, &m_0diff, 2);
        } else {
            if (t == last_gpstime[last]) { enc->encodeSymbol(&m_multi, (U32)G::MULTI_UNCHANGED); return; }
            const I64 d64 = (I64)(t - last_gpstime[last]);
            if (fits32(d64)) {
                const I32 d = (I32)d64, ld = last_gpstime_diff[last];
                const float mf = (float)d / (float)ld;
                const I32 multi = mf >= 0.0f ? (I32)(mf + 0.5f) : (I32)(mf - 0.5f);
                if (multi == 1) {
                    enc->encodeSymbol(&m_multi, 1);
                    ic.compress(ld, d, 1);
                    multi_extreme_counter[last] = 0;
                } else if (multi > 0) {
                    if (multi < G::MULTI) {
                        enc->encodeSymbol(&m_multi, (U32)multi);
                        ic.compress((I32)((U32)multi * (U32)ld), d, multi < 10 ? 2 : 3);
                    } else {
                        enc->encodeSymbol(&m_multi, (U32)G::MULTI);
                        ic.compress((I32)((U32)G::MULTI * (U32)ld), d, 4);
                        bump(d);
                    }
                } else if (multi < 0) {
                    if (multi > G::MULTI_MINUS) {
                        enc->encodeSymbol(&m_multi, (U32)(G::MULTI - multi));
                        ic.compress((I32)((U32)multi * (U32)ld), d, 5);
                    } else {
                        enc->encodeSymbol(&m_multi, (U32)(G::MULTI - G::MULTI_MINUS));
                        ic.compress((I32)((U32)G::MULTI_MINUS * (U32)ld), d, 6);
                        bump(d);
                    }
                } else {
                    enc->encodeSymbol(&m_multi, 0);
                    ic.compress(0, d, 7);
                    bump(d);
                }
                last_gpstime[last] = t;
            } else far_away(t, &m_multi, (U32)G::MULTI_CODE_FULL);
        }
    }
    void write(const U8* item) override { U64 t; memcpy(&t, item, 8); write_time(t); }
};

struct Rgb12Writer : ItemWriter {
    Encoder* enc;
    U16 last_item[3];
    SymbolModel m_byte_used, m_diff[6];
    explicit Rgb12Writer(Encoder* e) : enc(e) { m_byte_used.create(128); for (auto& m : m_diff) m.create(256); }
    void init(const U8* item) override { m_byte_used.init(); for (auto& m : m_diff) m.init(); memcpy(last_item, item, 6); }
    void write(const U8* in) override {
        U16 item[3];
        memcpy(item, in, 6);
        I32 diff_l = 0, diff_h = 0, corr;
        U32 sym = (U32)((last_item[0] & 0x00FF) != (item[0] & 0x00FF)) << 0;
        sym |= (U32)((last_item[0] & 0xFF00) != (item[0] & 0xFF00)) << 1;
        sym |= (U32)((last_item[1] & 0x00FF) != (item[1] & 0x00FF)) << 2;
        sym |= (U32)((last_item[1] & 0xFF00) != (item[1] & 0xFF00)) << 3;
        sym |= (U32)((last_item[2] & 0x00FF) != (item[2] & 0x00FF)) << 4;
        sym |= (U32)((last_item[2] & 0xFF00) != (item[2] & 0xFF00)) << 5;
        sym |= (U32)(((item[0] & 0x00FF) != (item[1] & 0x00FF)) || ((item[0] & 0x00FF) != (item[2] & 0x00FF)) ||
                     ((item[0] & 0xFF00) != (item[1] & 0xFF00)) || ((item[0] & 0xFF00) != (item[2] & 0xFF00))) << 6;
        enc->encodeSymbol(&m_byte_used, sym);
        if (sym & 1) { diff_l = (I32)(item[0] & 255) - (I32)(last_item[0] & 255); enc->encodeSymbol(&m_diff[0], u8_fold(diff_l)); }
        if (sym & 2) { diff_h = (I32)(item[0] >> 8) - (I32)(last_item[0] >> 8); enc->encodeSymbol(&m_diff[1], u8_fold(diff_h)); }
        if (sym & 64) {
            if (sym & 4) { corr = (I32)(item[1] & 255) - (I32)u8_clamp(diff_l + (last_item[1] & 255)); enc->encodeSymbol(&m_diff[2], u8_fold(corr)); }
            if (sym & 16) {
                diff_l = (diff_l + (I32)(item[1] & 255) - (I32)(last_item[1] & 255)) / 2;
                corr = (I32)(item[2] & 255) - (I32)u8_clamp(diff_l + (last_item[2] & 255));
                enc->encodeSymbol(&m_diff[4], u8_fold(corr));
            }
            if (sym & 8) { corr = (I32)(item[1] >> 8) - (I32)u8_clamp(diff_h + (last_item[1] >> 8)); enc->encodeSymbol(&m_diff[3], u8_fold(corr)); }
            if (sym & 32) {
                diff_h = (diff_h + (I32)(item[1] >> 8) - (I32)(last_item[1] >> 8)) / 2;
                corr = (I32)(item[2] >> 8) - (I32)u8_clamp(diff_h + (last_item[2] >> 8));
                enc->encodeSymbol(&m_diff[5], u8_fold(corr));
            }
        }
        memcpy(last_item, item, 6);
    }
};

struct ByteWriter : ItemWriter {
    Encoder* enc;
    U32 number;
    std::vector<U8> last_item;
    std::vector<SymbolModel> m_byte;
    ByteWriter(Encoder* e, U32 n) : enc(e), number(n), last_item(n), m_byte(n) { for (auto& m : m_byte) m.create(256); }
    void init(const U8* item) override { for (auto& m : m_byte) m.init(); memcpy(last_item.data(), item, number); }
    void write(const U8* item) override {
        for (U32 i = 0; i < number; ++i) enc->encodeSymbol(&m_byte[i], u8_fold((I32)item[i] - (I32)last_item[i]));
        memcpy(last_item.data(), item, number);
    }
};

}  // namespace laz
}  // namespace upcp

using namespace upcp;

extern "C" {

/* see include/upcp_cuda.h */
int upcp_laz_decode(const uint8_t* h_vlr, int32_t vlr_bytes, const uint8_t* h_data, int64_t data_bytes, int64_t data_file_offset,
                    int64_t n_points, int32_t point_size, uint8_t* h_out) {
    using namespace upcp::laz;
    if (!h_vlr || !h_data || !h_out || vlr_bytes < 34 || n_points < 0 || point_size < 20 || data_bytes < 8)
        UPCP_FAIL(UPCP_E_INVALID, "laz_decode: bad argument");
    U16 compressor, coder, num_items;
    I32 chunk_size;
    memcpy(&compressor, h_vlr, 2); memcpy(&coder, h_vlr + 2, 2); memcpy(&chunk_size, h_vlr + 12, 4); memcpy(&num_items, h_vlr + 32, 2);
    if (coder != 0) UPCP_FAIL(UPCP_E_CAPACITY, "laz_decode: unknown entropy coder");
    if (compressor != 2) UPCP_FAIL(UPCP_E_CAPACITY, "laz_decode: only the pointwise chunked compressor (LAS 1.0 - 1.3 point formats 0 - 3) is supported");
    if (vlr_bytes < 34 + 6 * (int)num_items) UPCP_FAIL(UPCP_E_INVALID, "laz_decode: truncated laszip VLR");
    if (chunk_size < 0) UPCP_FAIL(UPCP_E_CAPACITY, "laz_decode: variable-size chunks (chunk size 0xFFFFFFFF) are not supported");
    if (n_points == 0) return UPCP_OK;
    try {
        struct Spec { U16 type, size; };
        std::vector<Spec> specs;
        U32 total = 0;
        for (int i = 0; i < (int)num_items; ++i) {
            U16 type, size, version;
            memcpy(&type, h_vlr + 34 + 6 * i, 2); memcpy(&size, h_vlr + 36 + 6 * i, 2); memcpy(&version, h_vlr + 38 + 6 * i, 2);
            if (version != 2) UPCP_FAIL(UPCP_E_CAPACITY, "laz_decode: only version-2 item codecs are supported");
            if (!((type == 6 && size == 20) || (type == 7 && size == 8) || (type == 8 && size == 6) || (type == 0 && size >= 1)))
                UPCP_FAIL(UPCP_E_CAPACITY, "laz_decode: unsupported item type (wave packets / LAS 1.4 layered items)");
            specs.push_back({type, size});
            total += size;
        }
        if ((I32)total != point_size) UPCP_FAIL(UPCP_E_INVALID, "laz_decode: item sizes do not add up to the point size");
        // one decoder + item readers per worker (a chunk is an independent stream)
        struct Worker {
            Decoder dec;
            std::vector<ItemReader*> readers;
            explicit Worker(const std::vector<Spec>& sp) {
                for (const Spec& it : sp) {
                    if (it.type == 6) readers.push_back(new Point10Reader(&dec));
                    else if (it.type == 7) readers.push_back(new GpsTime11Reader(&dec));
                    else if (it.type == 8) readers.push_back(new Rgb12Reader(&dec));
                    else readers.push_back(new ByteReader(&dec, it.size));
                }
            }
            ~Worker() { for (auto p : readers) delete p; }
        };
        // The point data starts with the file position of the chunk table (the byte size of every chunk,
        // itself arithmetic-coded).  A sequential read does not need it -- every chunk ends where the decoder
        // stopped reading, LASzip relies on the same property -- but when it is there the chunk starts are
        // taken from it, as LASzip's integrity check does, and the chunks are decoded in parallel.
        const I64 per_chunk = chunk_size > 0 ? (I64)chunk_size : n_points;
        const I64 n_chunks = (n_points + per_chunk - 1) / per_chunk;
        std::vector<I64> chunk_start;                                  // offsets into h_data, empty = sequential
        {
            I64 table_pos;
            memcpy(&table_pos, h_data, 8);
            if (table_pos == -1 && data_bytes >= 16) memcpy(&table_pos, h_data + data_bytes - 8, 8);
            const I64 rel = table_pos - data_file_offset;
            if (table_pos > 0 && rel >= 8 && rel + 8 <= data_bytes) {
                U32 version, number_chunks;
                memcpy(&version, h_data + rel, 4); memcpy(&number_chunks, h_data + rel + 4, 4);
                if (version == 0 && (I64)number_chunks == n_chunks) {
                    ByteStream ts{h_data + rel + 8, h_data + data_bytes, false};
                    Decoder td;
                    td.init(&ts);
                    IntegerCompressor ic;
                    ic.create(&td, 32, 2);
                    ic.init();
                    I64 pos = 8;
                    I32 prev = 0;
                    for (U32 c = 0; c < number_chunks; ++c) {
                        chunk_start.push_back(pos);
                        prev = ic.decompress(c ? prev : 0, 1);
                        pos += (I64)(U32)prev;
                    }
                    if (ts.overrun || pos != rel) chunk_start.clear();       // inconsistent table: read sequentially
                    else chunk_start.push_back(rel);                         // (end of the last chunk)
                }
            }
        }
        // points [c * per_chunk, ...) of chunk c from `in`; false = truncated / overrun
        auto decode_chunk = [&](Worker& w, ByteStream& in, I64 c) -> bool {
            const I64 i0 = c * per_chunk, i1 = (i0 + per_chunk < n_points) ? i0 + per_chunk : n_points;
            for (I64 i = i0; i < i1; ++i) {
                U8* out = h_out + i * point_size;
                U32 off = 0;
                if (i == i0) {
                    // first point of a chunk: stored raw; then the coder starts afresh
                    if (in.end - in.p < point_size) return false;
                    memcpy(out, in.p, point_size);
                    in.p += point_size;
                    for (size_t r = 0; r < w.readers.size(); ++r) { w.readers[r]->init(out + off); off += specs[r].size; }
                    if (i + 1 < i1) w.dec.init(&in);
                } else {
                    for (size_t r = 0; r < w.readers.size(); ++r) { w.readers[r]->read(out + off); off += specs[r].size; }
                }
                if (in.overrun) return false;
            }
            return true;
        };
        bool ok = true;
        if (chunk_start.empty()) {
            Worker w(specs);
            ByteStream in{h_data + 8, h_data + data_bytes, false};
            for (I64 c = 0; c < n_chunks && ok; ++c) ok = decode_chunk(w, in, c);
        } else {
            unsigned hw = std::thread::hardware_concurrency();
            const I64 n_threads = std::max<I64>(1, std::min<I64>(std::min<I64>(hw ? hw : 1, 32), n_chunks));
            std::atomic<I64> next{0};
            std::atomic<int> failed{0}, oom{0};
            auto body = [&]() {
                try {
                    Worker w(specs);
                    for (I64 c = next.fetch_add(1); c < n_chunks && !failed.load(); c = next.fetch_add(1)) {
                        ByteStream in{h_data + chunk_start[(size_t)c], h_data + data_bytes, false};
                        if (!decode_chunk(w, in, c)) failed.store(1);
                    }
                } catch (const std::bad_alloc&) { oom.store(1); failed.store(1); }
            };
            std::vector<std::thread> pool;
            for (I64 t = 1; t < n_threads; ++t) pool.emplace_back(body);
            body();
            for (auto& t : pool) t.join();
            if (oom.load()) UPCP_FAIL(UPCP_E_INTERNAL, "laz_decode: out of host memory");
            ok = !failed.load();
        }
        if (!ok) UPCP_FAIL(UPCP_E_INVALID, "laz_decode: compressed stream ended early");
    } catch (const std::bad_alloc&) {
        UPCP_FAIL(UPCP_E_INTERNAL, "laz_decode: out of host memory");
    }
    return UPCP_OK;
}

/* see include/upcp_cuda.h */
int upcp_laz_encode(const uint8_t* h_points, int64_t n_points, int32_t point_size, int32_t point_format, int32_t chunk_size,
                    int64_t data_file_offset, uint8_t* h_vlr58, int32_t* vlr_bytes, uint8_t* h_out, int64_t out_cap,
                    int64_t* out_bytes) {
    using namespace upcp::laz;
    if (!h_vlr58 || !vlr_bytes || !out_bytes || n_points < 0 || (n_points > 0 && !h_points) || chunk_size < 1 || data_file_offset < 0)
        UPCP_FAIL(UPCP_E_INVALID, "laz_encode: bad argument");
    *out_bytes = 0;
    static const int kBase[4] = {20, 28, 26, 34};
    if (point_format < 0 || point_format > 3) UPCP_FAIL(UPCP_E_CAPACITY, "laz_encode: point formats 0 - 3 only (LAS 1.4 formats need the layered codecs)");
    const int extra = point_size - kBase[point_format];
    if (extra < 0 || extra > 65535) UPCP_FAIL(UPCP_E_INVALID, "laz_encode: point size smaller than the point format's record");
    try {
        // items of the record: POINT10 [GPSTIME11] [RGB12] [BYTE x extra], all version 2
        struct Item { U16 type, size, version; };
        std::vector<Item> items;
        items.push_back({6, 20, 2});
        if (point_format == 1 || point_format == 3) items.push_back({7, 8, 2});
        if (point_format == 2 || point_format == 3) items.push_back({8, 6, 2});
        if (extra > 0) items.push_back({0, (U16)extra, 2});
        // the "laszip encoded" VLR payload (record id 22204)
        U8* v = h_vlr58;
        memset(v, 0, 58);
        const U16 compressor = 2, coder = 0, revision = 0, num_items = (U16)items.size();
        const U32 options = 0;
        const I64 minus1 = -1;
        memcpy(v, &compressor, 2); memcpy(v + 2, &coder, 2);
        v[4] = 2; v[5] = 2;                                   // the format version of LASzip 2.2: chunked, version-2 items
        memcpy(v + 6, &revision, 2); memcpy(v + 8, &options, 4); memcpy(v + 12, &chunk_size, 4);
        memcpy(v + 16, &minus1, 8); memcpy(v + 24, &minus1, 8); memcpy(v + 32, &num_items, 2);
        for (size_t i = 0; i < items.size(); ++i) {
            memcpy(v + 34 + 6 * i, &items[i].type, 2); memcpy(v + 36 + 6 * i, &items[i].size, 2); memcpy(v + 38 + 6 * i, &items[i].version, 2);
        }
        *vlr_bytes = 34 + 6 * (int)items.size();

        struct Worker {
            Encoder enc;
            std::vector<ItemWriter*> writers;
            explicit Worker(const std::vector<Item>& its) {
                for (const Item& it : its) {
                    if (it.type == 6) writers.push_back(new Point10Writer(&enc));
                    else if (it.type == 7) writers.push_back(new GpsTime11Writer(&enc));
                    else if (it.type == 8) writers.push_back(new Rgb12Writer(&enc));
                    else writers.push_back(new ByteWriter(&enc, it.size));
                }
            }
            ~Worker() { for (auto p : writers) delete p; }
        };
        // every chunk is an independent stream: first point raw, the rest through a freshly initialised coder;
        // the chunks are encoded in parallel and laid out in order
        const I64 n_chunks = (n_points + chunk_size - 1) / chunk_size;
        std::vector<std::vector<U8>> parts((size_t)n_chunks);
        auto encode_chunk = [&](Worker& w, I64 c) {
            std::vector<U8> o;                       // (local: neighbouring entries of `parts` share cache lines)
            const I64 i0 = c * (I64)chunk_size, i1 = (i0 + chunk_size < n_points) ? i0 + chunk_size : n_points;
            o.reserve((size_t)((i1 - i0) * point_size / 3 + point_size + 16));
            for (I64 i = i0; i < i1; ++i) {
                const U8* pt = h_points + i * point_size;
                U32 off = 0;
                if (i == i0) {
                    o.insert(o.end(), pt, pt + point_size);
                    for (size_t k = 0; k < w.writers.size(); ++k) { w.writers[k]->init(pt + off); off += items[k].size; }
                    w.enc.init(&o);
                } else {
                    for (size_t k = 0; k < w.writers.size(); ++k) { w.writers[k]->write(pt + off); off += items[k].size; }
                }
            }
            w.enc.done();
            parts[(size_t)c] = std::move(o);
        };
        {
            unsigned hw = std::thread::hardware_concurrency();
            const I64 n_threads = std::max<I64>(1, std::min<I64>(std::min<I64>(hw ? hw : 1, 32), n_chunks));
            std::atomic<I64> next{0};
            std::atomic<int> oom{0};
            auto body = [&]() {
                try {
                    Worker w(items);
                    for (I64 c = next.fetch_add(1); c < n_chunks && !oom.load(); c = next.fetch_add(1)) encode_chunk(w, c);
                } catch (const std::bad_alloc&) { oom.store(1); }
            };
            std::vector<std::thread> pool;
            for (I64 t = 1; t < n_threads; ++t) pool.emplace_back(body);
            if (n_chunks > 0) body();
            for (auto& t : pool) t.join();
            if (oom.load()) UPCP_FAIL(UPCP_E_INTERNAL, "laz_encode: out of host memory");
        }
        std::vector<U8> out(8, 0);                            // file position of the chunk table, patched below
        std::vector<U32> chunk_bytes;
        for (const auto& part : parts) { out.insert(out.end(), part.begin(), part.end()); chunk_bytes.push_back((U32)part.size()); }
        parts.clear();
        Encoder enc;
        // chunk table: version, count, then the chunk sizes through the integer compressor (context 1)
        const I64 table_pos = data_file_offset + (I64)out.size();
        memcpy(out.data(), &table_pos, 8);
        const U32 version = 0, number_chunks = (U32)chunk_bytes.size();
        U8 head[8];
        memcpy(head, &version, 4); memcpy(head + 4, &number_chunks, 4);
        out.insert(out.end(), head, head + 8);
        if (number_chunks) {
            enc.init(&out);
            IntegerEncoder ic;
            ic.create(&enc, 32, 2);
            ic.init();
            for (U32 c = 0; c < number_chunks; ++c) ic.compress(c ? (I32)chunk_bytes[c - 1] : 0, (I32)chunk_bytes[c], 1);
            enc.done();
        }
        *out_bytes = (int64_t)out.size();
        if (!h_out || out_cap < (int64_t)out.size()) UPCP_FAIL(UPCP_E_WORKSPACE, "laz_encode: output buffer too small (out_bytes holds the size)");
        memcpy(h_out, out.data(), out.size());
    } catch (const std::bad_alloc&) {
        UPCP_FAIL(UPCP_E_INTERNAL, "laz_encode: out of host memory");
    }
    return UPCP_OK;
}

}  // extern "C"
