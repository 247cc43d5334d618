// Host-side ranging-code generation for BDS-3 B1C (Weil codes) and B2a (Gold-type
// LFSR codes).  Integer, table driven.  Reference behaviour:
//   BDS-3_B1C/include/generateDataBOC11.m:43-90, generatePilotBOC11.m:44-94,
//   generatePilotBOC61.m:44-96, JacobiSymbol.m:48-123
//   BDS-3_B2a/include/generateB2aDataCode.m:39-138, generateB2aPilotCode.m:39-138
#pragma once
#include <cstdint>
#include <vector>

namespace bds {

constexpr int kCodeLen = 10230;
constexpr int kWeilN = 10243;

// Primary codes as 0/1 chips (1 == bipolar -1), prn in 1..63.
// component: BDS_CODE_B1C_DATA_PRIMARY, BDS_CODE_B1C_PILOT_PRIMARY, BDS_CODE_B2A_DATA, BDS_CODE_B2A_PILOT
bool primary_bits(int component, int prn, std::vector<uint8_t>& chips);

// Full component as +1/-1 int8 (including BOC sub-carrier expansion).
int component_length(int component);
bool gen_component(int component, int prn, std::vector<int8_t>& out);

// Bit-packed primary code, little-endian bit order inside 32-bit words
// (bit k of word w = chip 32*w+k), padded to kPackedWords words.
constexpr int kPackedWords = 320;  // 10240 bits >= 10230
void pack_bits(const std::vector<uint8_t>& chips, uint32_t* words);

}  // namespace bds
