// ugbase/common/util/crc32.cpp:34 includes boost/crc.hpp (boost is not part of /root/reference); ug::crc32 is only
// used to hash debug-id names (common/debug_id.cpp).  Any deterministic hash serves the test binary.
#include "common/types.h"
namespace ug {
uint32 crc32(const char* s)
{
	uint32 h = 2166136261u;
	for (; s && *s; ++s) { h ^= (unsigned char)*s; h *= 16777619u; }
	return h;
}
} // namespace ug
