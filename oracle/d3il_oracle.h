/* d3il_oracle.h — C interface of the fp64 CPU oracle (test infrastructure; see d3il_oracle.c header). */
#ifndef D3IL_ORACLE_H
#define D3IL_ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct Env Env;
Env* d3o_create(const void* scene_blob, size_t nbytes);
void d3o_destroy(Env* e);
const char* d3o_last_error(void);
void d3o_reset(Env* e, const double* ctx /* [nobj*7] xyz+quat per free object, or NULL */);
void d3o_step(Env* e, const double* action, float* obs, double* reward, int* done, double* info);
void d3o_substep(Env* e, int n);
void d3o_robot_state(const Env* e, double* tcp3);
void d3o_joint_state(const Env* e, double* j8);
void d3o_get_obs(const Env* e, float* obs);
int d3o_state_dim(const Env* e);
void d3o_get_state(const Env* e, double* out);
void d3o_set_state(Env* e, const double* in);
/* probes used by the invariant tests */
void d3o_forward(Env* e);
int d3o_probe(const Env* e, const char* what, double* out, int cap);
void d3o_ik_fk(Env* e, const double* q, double* pos, double* quat, double* J);
int d3o_collide(int t1, const double* p1, const double* q1, const double* s1, int t2, const double* p2, const double* q2, const double* s2, double* out);
#ifdef __cplusplus
}
#endif
#endif
