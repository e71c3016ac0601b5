// ministream stand-in for <Pothos/Config.hpp>: export/visibility macros only.
#pragma once
#ifndef POTHOS_API
#define POTHOS_API
#endif
#ifndef POTHOS_LOCAL
#define POTHOS_LOCAL
#endif
