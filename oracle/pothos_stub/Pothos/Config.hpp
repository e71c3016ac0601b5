// Empty stand-in for <Pothos/Config.hpp> (ChirpGenerator.hpp:5 includes it
// only for export macros it never uses).  Ours; test infrastructure only.
#pragma once
