// See bds_codes.h.  The (w, p) and register-2 tables are ICD constants that the
// reference carries as literals (file:line in the header).
#include "bds_codes.h"

#include <cstring>
#include <mutex>

#include "../../include/bdsgpu.h"

namespace bds {
namespace {

const uint16_t kDataW[63] = {2678, 4802, 958,  859,  3843, 2232, 124,  4352, 1816, 1126, 1860, 4800, 2267,
                             424,  4192, 4333, 2656, 4148, 243,  1330, 1593, 1470, 882,  3202, 5095, 2546,
                             1733, 4795, 4577, 1627, 3638, 2553, 3646, 1087, 1843, 216,  2245, 726,  1966,
                             670,  4130, 53,   4830, 182,  2181, 2006, 1080, 2288, 2027, 271,  915,  497,
                             139,  3693, 2054, 4342, 3342, 2592, 1007, 310,  4203, 455,  4318};
const uint16_t kDataP[63] = {699,  694,  7318, 2127, 715,  6682, 7850, 5495, 1162, 7682, 6792, 9973, 6596,
                             2092, 19,   10151, 6297, 5766, 2359, 7136, 1706, 2128, 6827, 693,  9729, 1620,
                             6805, 534,  712,  1929, 5355, 6139, 6339, 1470, 6867, 7851, 1162, 7659, 1156,
                             2672, 6043, 2862, 180,  2663, 6940, 1645, 1582, 951,  6878, 7701, 1823, 2391,
                             2606, 822,  6403, 239,  442,  6769, 2560, 2502, 5072, 7268, 341};
const uint16_t kPilotW[63] = {796,  156,  4198, 3941, 1374, 1338, 1833, 2521, 3175, 168,  2715, 4408, 3160,
                              2796, 459,  3594, 4813, 586,  1428, 2371, 2285, 3377, 4965, 3779, 4547, 1646,
                              1430, 607,  2118, 4709, 1149, 3283, 2473, 1006, 3670, 1817, 771,  2173, 740,
                              1433, 2458, 3459, 2155, 1205, 413,  874,  2463, 1106, 1590, 3873, 4026, 4272,
                              3556, 128,  1200, 130,  4494, 1871, 3073, 4386, 4098, 1923, 1176};
const uint16_t kPilotP[63] = {7575, 2369, 5688, 539,  2270, 7306, 6457, 6254, 5644, 7119, 1402, 5557, 5764,
                              1073, 7001, 5910, 10060, 2710, 1546, 6887, 1883, 5613, 5062, 1038, 10170, 6484,
                              1718, 2535, 1158, 526,  7331, 5844, 6423, 6968, 1280, 1838, 1989, 6468, 2091,
                              1581, 1453, 6252, 7122, 7711, 7216, 2113, 1095, 1628, 1713, 6102, 6123, 6070,
                              1115, 8047, 6795, 2575, 53,   1729, 6388, 682,  5565, 7160, 2277};
// register-2 start states: bit k = stage k+1
const uint16_t kB2aG2[63] = {0x1481, 0x581,  0x16a1, 0x1e51, 0x1551, 0xeb1,  0xef1,  0x1bf1, 0x1299, 0xb79,  0x1585,
                             0x445,  0x1545, 0x1b45, 0x745,  0x18a5, 0x1de5, 0x1015, 0xf95,  0x1ab5, 0x11b5, 0x194d,
                             0x8cd,  0x32d,  0xdad,  0x9ed,  0x1fed, 0x91d,  0x79d,  0x10bd, 0x27d,  0x57d,  0x1afd,
                             0x19fd, 0x1143, 0x523,  0x1da3, 0x1113, 0x1313, 0x1ab3, 0x11b3, 0x973,  0x154b, 0x5cb,
                             0x1a6b, 0x1d5b, 0x587,  0x1827, 0x1a27, 0x18a7, 0x2a7,  0x1b97, 0x1d37, 0x24f,  0x52f,
                             0x132f, 0xb6f,  0x3ef,  0x1fef, 0x15bf, 0x804,  0x15fb, 0x978};
const uint16_t kB2aPilotTail[3] = {0xc25, 0x3f4, 0x1558};  // PRN 61..63 differ for the pilot

std::once_flag g_leg_once;
uint8_t g_legendre[kWeilN];

void build_legendre() {
    // L(k) = 1 iff k is a non-zero square mod 10243 (prime): the reference's
    // JacobiSymbol(k, 10243) == +1 case; -1 and 0 map to 0.
    std::memset(g_legendre, 0, sizeof(g_legendre));
    for (long long k = 1; k < kWeilN; ++k) g_legendre[(k * k) % kWeilN] = 1;
}

void weil(int w, int p, std::vector<uint8_t>& chips) {
    std::call_once(g_leg_once, build_legendre);
    chips.resize(kCodeLen);
    for (int n = 0; n < kCodeLen; ++n) {
        int k = (n + p - 1) % kWeilN;
        chips[n] = g_legendre[k] ^ g_legendre[(k + w) % kWeilN];
    }
}

// Logic-level (0/1) form of the reference's +-1 registers: product of taps == XOR.
void b2a_lfsr(uint32_t g2, uint32_t taps1, uint32_t taps2, std::vector<uint8_t>& chips) {
    chips.resize(kCodeLen);
    uint32_t r1 = 0x1fff, r2 = g2 & 0x1fff;  // bit k = stage k+1
    for (int n = 1; n <= kCodeLen; ++n) {
        chips[n - 1] = ((r1 >> 12) ^ (r2 >> 12)) & 1u;
        uint32_t f1 = __builtin_parity(r1 & taps1);
        uint32_t f2 = __builtin_parity(r2 & taps2);
        r1 = ((r1 << 1) | f1) & 0x1fff;
        r2 = ((r2 << 1) | f2) & 0x1fff;
        if (n == 8190) r1 = 0x1fff;
    }
}

uint32_t tapmask(std::initializer_list<int> taps) {
    uint32_t m = 0;
    for (int t : taps) m |= 1u << (t - 1);
    return m;
}

}  // namespace

bool primary_bits(int component, int prn, std::vector<uint8_t>& chips) {
    if (prn < 1 || prn > 63) return false;
    switch (component) {
        case BDS_CODE_B1C_DATA_PRIMARY:
        case BDS_CODE_B1C_DATA_BOC11:
            weil(kDataW[prn - 1], kDataP[prn - 1], chips);
            return true;
        case BDS_CODE_B1C_PILOT_PRIMARY:
        case BDS_CODE_B1C_PILOT_BOC11:
        case BDS_CODE_B1C_PILOT_BOC61:
            weil(kPilotW[prn - 1], kPilotP[prn - 1], chips);
            return true;
        case BDS_CODE_B2A_DATA:
            b2a_lfsr(kB2aG2[prn - 1], tapmask({1, 5, 11, 13}), tapmask({3, 5, 9, 11, 12, 13}), chips);
            return true;
        case BDS_CODE_B2A_PILOT:
            b2a_lfsr(prn >= 61 ? kB2aPilotTail[prn - 61] : kB2aG2[prn - 1], tapmask({3, 6, 7, 13}),
                     tapmask({1, 5, 7, 8, 12, 13}), chips);
            return true;
        default:
            return false;
    }
}

int component_length(int component) {
    switch (component) {
        case BDS_CODE_B1C_DATA_PRIMARY:
        case BDS_CODE_B1C_PILOT_PRIMARY:
        case BDS_CODE_B2A_DATA:
        case BDS_CODE_B2A_PILOT:
            return kCodeLen;
        case BDS_CODE_B1C_DATA_BOC11:
        case BDS_CODE_B1C_PILOT_BOC11:
            return kCodeLen * 2;
        case BDS_CODE_B1C_PILOT_BOC61:
            return kCodeLen * 12;
        default:
            return -1;
    }
}

bool gen_component(int component, int prn, std::vector<int8_t>& out) {
    std::vector<uint8_t> chips;
    if (!primary_bits(component, prn, chips)) return false;
    int sub = component_length(component) / kCodeLen;
    out.resize((size_t)kCodeLen * sub);
    for (int c = 0; c < kCodeLen; ++c) {
        int8_t v = chips[c] ? -1 : 1;
        // sub-carrier: BOC(1,1) chip -> [-c, +c]; BOC(6,1) chip -> (-1)^ii c, ii = 1..12
        for (int i = 0; i < sub; ++i) out[(size_t)c * sub + i] = (sub == 1) ? v : ((i & 1) ? v : (int8_t)-v);
    }
    return true;
}

void pack_bits(const std::vector<uint8_t>& chips, uint32_t* words) {
    std::memset(words, 0, sizeof(uint32_t) * kPackedWords);
    for (size_t i = 0; i < chips.size(); ++i)
        if (chips[i]) words[i >> 5] |= 1u << (i & 31);
}

}  // namespace bds
