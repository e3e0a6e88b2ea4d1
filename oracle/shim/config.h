/* Build shim for compiling the UNMODIFIED reference out of /root/reference
 * without autotools (src/patternmodeller.cpp:19 includes "config.h").
 * Test infrastructure only. */
#ifndef ORACLE_SHIM_CONFIG_H
#define ORACLE_SHIM_CONFIG_H
#define VERSION "2.5.9"
#define PACKAGE_VERSION "2.5.9"
#endif
