/* Hand-written replacement for the autoconf-generated config.h of the
 * reference (autotools are not available in this image).  TEST
 * INFRASTRUCTURE ONLY: used when compiling the unmodified reference sources
 * from /root/reference into oracle/_ref/ (see oracle/Makefile).
 * The reference only needs ARCHNAME (initug.cc) plus the usual HAVE_* set. */
#ifndef ORACLE_SHIM_CONFIG_H
#define ORACLE_SHIM_CONFIG_H
#define ARCHNAME "x86_64-linux"
#define ARCH_CPU "x86_64"
#define ARCH_OS "linux-gnu"
#define ARCH_VENDOR "pc"
#define DYNAMIC_MEMORY_ALLOCMODEL 1
#define HAVE_STDLIB_H 1
#define HAVE_STRING_H 1
#define HAVE_UNISTD_H 1
#define HAVE_LIMITS_H 1
#define HAVE_FLOAT_H 1
#define HAVE_MALLOC_H 1
#define HAVE_VALUES_H 1
#define STDC_HEADERS 1
#define TIME_WITH_SYS_TIME 1
#define PACKAGE "ug"
#define VERSION "3.12.1"
#endif
