// what a REAL*8 dummy reads when it is handed a REAL(16) literal by reference (x86-64, little endian):
// the low 8 bytes of the binary128 value of the decimal string.  g++ tools/hrs_literal_probe.cpp -lquadmath
// The table in oracle/arms.cpp (hrs_lit) was printed by this program.
#include <quadmath.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
int main() {
  const char* lits[] = {"62.75333333", "31.37666667", "121.77333333", "60.88666667", "659.73445725", "121.7866667", "60.89333333", "45.0"};
  for (const char* l : lits) {
    const __float128 q = strtoflt128(l, nullptr);
    uint64_t lo;
    std::memcpy(&lo, &q, 8);
    double d;
    std::memcpy(&d, &lo, 8);
    std::printf("    {\"%s\", 0x%016llxULL},   // %.17g\n", l, (unsigned long long)lo, d);
  }
}
