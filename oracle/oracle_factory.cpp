// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_engine.hpp).
// Links the host driver against the CPU oracle engine -> oracle/_build/gamba_oracle.
#include "../gamba_b200/host/f4.hpp"
#include "oracle_engine.hpp"
#include <cstdlib>
// GAMBA_ORACLE_THREADS=N (N > 1): checker mode on N host threads (tests on the big named systems);
// default: the single-thread timing mode
namespace gb {
Engine* gb_make_engine(const Params&, uint32_t) {
    const char* t = std::getenv("GAMBA_ORACLE_THREADS");
    return new OracleEngine(t ? (unsigned)std::atoi(t) : 1u);
}
}
