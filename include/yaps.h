/*
 * yaps.h -- message / fatal-error helpers with a pluggable sink.
 * Same four entry points as the reference's lib/yaps.h:16-19, so programs written against
 * libstb (test/list.c, test/demo.c) link unchanged.
 */
#ifndef STB_B200_YAPS_H
#define STB_B200_YAPS_H
#include <stdarg.h>
#ifdef __cplusplus
extern "C" {
#endif
/* print through the sink (stderr by default) */
void yaps_message(const char *fmt, ...);
/* print, then exit(1) */
void yaps_quit(const char *fmt, ...);
/* print strerror(errno) + message, then exit(1) */
void yaps_sysquit(const char *fmt, ...);
/* install a sink; NULL restores stderr */
void yaps_yapper(void (*yapper)(const char *format, va_list ap));
#ifdef __cplusplus
}
#endif
#endif
