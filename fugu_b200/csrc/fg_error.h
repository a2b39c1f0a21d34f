// Shared error reporting of the library (thread-local message behind fg_last_error()).
#pragma once
#include <stdint.h>
namespace fg {
// records a printf-style message for fg_last_error() on this thread and returns `code`
int32_t host_fail(int32_t code, const char* fmt, ...);
}
