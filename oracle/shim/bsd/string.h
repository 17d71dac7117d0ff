/* libbsd is absent from this image; the reference (radio.c:10-12, modes.c:11-13) only needs strlcpy. */
#ifndef KA9Q_ORACLE_BSD_STRING_SHIM_H
#define KA9Q_ORACLE_BSD_STRING_SHIM_H 1
#include <stddef.h>
#include <string.h>
size_t strlcpy(char *dst, const char *src, size_t size);
#endif
