/*
 * TEST INFRASTRUCTURE ONLY (oracle build).  Minimal stand-in for <rpc/types.h>:
 * this image has no libtirpc, and the reference includes <rpc/types.h> and
 * <rpc/xdr.h> (kd.c:7-8, totipnat.c:3-4).  Only what those call sites need.
 */
#ifndef SKID_SHIM_RPC_TYPES_H
#define SKID_SHIM_RPC_TYPES_H

#include <stdio.h>
#include <stdint.h>
#include <string.h>

typedef int bool_t;
typedef unsigned int u_int;
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif

#endif
