/* Minimal XDR stand-in: glibc no longer ships <rpc/rpc.h> and libtirpc is not
 * installed.  The reference includes it from low/bio.cc and ui/fieldio.cc for
 * binary grid I/O, which the oracle never calls.  TEST INFRASTRUCTURE ONLY.
 * The functions are defined in oracle/shim/xdr_stub.c and always fail. */
#ifndef ORACLE_SHIM_RPC_H
#define ORACLE_SHIM_RPC_H
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
enum xdr_op { XDR_ENCODE = 0, XDR_DECODE = 1, XDR_FREE = 2 };
typedef struct { enum xdr_op x_op; FILE *x_file; } XDR;
typedef int bool_t;
void xdrstdio_create(XDR *xdrs, FILE *file, enum xdr_op op);
bool_t xdr_int(XDR *xdrs, int *ip);
bool_t xdr_u_int(XDR *xdrs, unsigned int *up);
bool_t xdr_double(XDR *xdrs, double *dp);
bool_t xdr_float(XDR *xdrs, float *fp);
bool_t xdr_char(XDR *xdrs, char *cp);
#define xdr_destroy(x) ((void)0)
#ifdef __cplusplus
}
#endif
#endif
