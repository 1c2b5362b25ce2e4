// Stand-in for the CMake-generated chase_config.h of the reference
// (/root/reference/chase_config.h.in); only the version string is needed.
#pragma once
#define CHASE_VERSION "1.7.0"
