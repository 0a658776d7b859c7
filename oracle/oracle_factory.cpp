// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_engine.hpp).
// Links the host driver against the CPU oracle engine -> oracle/_build/gamba_oracle.
#include "../gamba_b200/host/f4.hpp"
#include "oracle_engine.hpp"
namespace gb { Engine* gb_make_engine(const Params&, uint32_t) { return new OracleEngine(); } }
