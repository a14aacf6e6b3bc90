/* See rpc/rpc.h in this directory.  Every call reports failure. */
#include "rpc/rpc.h"
void xdrstdio_create(XDR *x, FILE *f, enum xdr_op op) { x->x_op = op; x->x_file = f; }
bool_t xdr_int(XDR *x, int *p) { (void)x; (void)p; return 0; }
bool_t xdr_u_int(XDR *x, unsigned int *p) { (void)x; (void)p; return 0; }
bool_t xdr_double(XDR *x, double *p) { (void)x; (void)p; return 0; }
bool_t xdr_float(XDR *x, float *p) { (void)x; (void)p; return 0; }
bool_t xdr_char(XDR *x, char *p) { (void)x; (void)p; return 0; }
