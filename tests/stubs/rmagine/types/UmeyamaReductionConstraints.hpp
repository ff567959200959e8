// STUB of rmagine/types/UmeyamaReductionConstraints.hpp (tests/stubs/README.md)
#pragma once
namespace rmagine { struct UmeyamaReductionConstraints { float max_dist; }; }
