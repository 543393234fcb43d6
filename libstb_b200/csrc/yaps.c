/*
 * yaps.c -- message / fatal-error helpers (interface of the reference's lib/yaps.c:33-81).
 * The library's error convention lives here: anything unrecoverable prints and exit(1)s.
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "yaps.h"

static void (*sink)(const char *format, va_list ap) = NULL;

void yaps_yapper(void (*yapper)(const char *format, va_list ap)) { sink = yapper; }

static void emit(const char *fmt, va_list ap) {
  if (sink)
    sink(fmt, ap);
  else
    vfprintf(stderr, fmt, ap);
}

static void emit_str(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  emit(fmt, ap);
  va_end(ap);
}

void yaps_message(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  emit(fmt, ap);
  va_end(ap);
}

void yaps_quit(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  emit(fmt, ap);
  va_end(ap);
  exit(1);
}

void yaps_sysquit(const char *fmt, ...) {
  va_list ap;
  emit_str("%s: ", strerror(errno));
  va_start(ap, fmt);
  emit(fmt, ap);
  va_end(ap);
  exit(1);
}
