/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (timed CPU baseline of bench.py; never linked by the product library).
 *
 * The whole `VecTask.step` of the flat Anymal / Hound task for a batch of environments on the host cores: the float32
 * instantiation of the dynamics restatement (oracle_dyn_impl.h) followed by the task's post_physics_step, one OpenMP
 * loop over the environments (they are independent).  This is what `bench.py --impl reference` and the `cpu_baseline`
 * leg time (kind "port": Isaac Gym / PhysX, the reference's real CPU pipeline, is a closed binary that is not installed).
 * It is compiled ON THE MACHINE THAT RUNS IT with -O3 -march=native -fopenmp (oracle/cpu_baseline.py::build_native).
 *
 * Reference order restated: vec_task.py:374 (clamp actions) -> tasks/anymal.py:226-229 (position targets) ->
 * vec_task.py:379-382 (simulate) -> tasks/anymal.py:231-239 (progress, reset_idx :278-304, observations :354-386,
 * reward :311-351) -> vec_task.py:394,402 (time-outs, clamp obs).  Checked against the numpy composition
 * (oracle/cpu_baseline.py::CpuAnymalStep) in tests/test_cpu_baseline.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include "b200gym.h"

#define ORC_REAL float
#define ORC_SUF _sf32
#include "oracle_dyn_impl.h"
#undef ORC_REAL
#undef ORC_SUF

#ifdef _OPENMP
#include <omp.h>
#endif

static void rot_inv(const float* q, const float* v, float* o) {     /* utils/torch_jit_utils.py:93-103 */
    float w = q[3], s = 2.0f * w * w - 1.0f;
    float c[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
    float d = 2.0f * (q[0] * v[0] + q[1] * v[1] + q[2] * v[2]);
    for (int i = 0; i < 3; i++) o[i] = v[i] * s - c[i] * w * 2.0f + q[i] * d;
}
static void rot_fwd(const float* q, const float* v, float* o) {     /* utils/torch_jit_utils.py:80-90 */
    float w = q[3], s = 2.0f * w * w - 1.0f;
    float c[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
    float d = 2.0f * (q[0] * v[0] + q[1] * v[1] + q[2] * v[2]);
    for (int i = 0; i < 3; i++) o[i] = v[i] * s + c[i] * w * 2.0f + q[i] * d;
}
static float urange(float lo, float hi, float u) { return (hi - lo) * u + lo; }   /* torch_rand_float :215-218 */

int orc_omp_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* draws: (N, 2 nd + 3) uniforms consumed by the environments that reset (dof pos, dof vel, cmd x, y, yaw) */
int orc_anymal_step_omp(const b2g_model* m, const b2g_sim_params* sp, const b2g_dof_props* dp, const b2g_anymal_cfg* c, int n_envs, int threads,
                        float* root, float* dof, const float* actions_in, const float* draws, float* commands, int64_t* progress, int64_t* reset,
                        float* torques, float* contact, float* obs, float* obs_clamped, float* rew, int64_t* timeout) {
    const int nd = m->n_dof, nb = m->n_bodies, no = 12 + 3 * nd, ndraw = 2 * nd + 3;
    const int nsub = sp->substeps > 0 ? sp->substeps : 1;
    const float h = sp->dt / (float)nsub;
    int rc = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static) reduction(| : rc)
    for (int e = 0; e < n_envs; e++) {
        float* r = root + (size_t)e * 13;
        float* d = dof + (size_t)e * nd * 2;
        float act[B2G_MAX_DOF], tgt[B2G_MAX_DOF], zero[B2G_MAX_DOF];
        for (int j = 0; j < nd; j++) {
            float a = actions_in[(size_t)e * nd + j];
            a = a < -c->clip_actions ? -c->clip_actions : (a > c->clip_actions ? c->clip_actions : a);
            act[j] = a;
            tgt[j] = c->action_scale * a + c->default_dof_pos[j];
            zero[j] = 0.0f;
        }
        for (int s = 0; s < nsub; s++)
            if (orc_substep_sf32(m, sp, dp, 0, 0, 1.0f, h, r, d, tgt, zero, torques + (size_t)e * nd, contact + (size_t)e * nb * 3) != 0) rc |= 1;
        /* post_physics_step */
        progress[e] += 1;
        float* cmd = commands + (size_t)e * 3;
        if (reset[e] != 0) {
            const float* u = draws + (size_t)e * ndraw;
            for (int j = 0; j < nd; j++) {
                d[2 * j] = c->default_dof_pos[j] * urange(0.5f, 1.5f, u[j]);
                d[2 * j + 1] = urange(-0.1f, 0.1f, u[nd + j]);
            }
            for (int k = 0; k < 13; k++) r[k] = c->init_root[k];
            cmd[0] = urange(c->cmd_x[0], c->cmd_x[1], u[2 * nd]);
            cmd[1] = urange(c->cmd_y[0], c->cmd_y[1], u[2 * nd + 1]);
            cmd[2] = urange(c->cmd_yaw[0], c->cmd_yaw[1], u[2 * nd + 2]);
            progress[e] = 0;
        }
        float lin[3], ang[3], pg[3];
        const float gdir[3] = {0.0f, 0.0f, -1.0f};
        rot_inv(r + 3, r + 7, lin);
        rot_inv(r + 3, r + 10, ang);
        rot_fwd(r + 3, gdir, pg);
        float* o = obs + (size_t)e * no;
        for (int k = 0; k < 3; k++) { o[k] = lin[k] * c->lin_vel_scale; o[3 + k] = ang[k] * c->ang_vel_scale; o[6 + k] = pg[k]; }
        o[9] = cmd[0] * c->lin_vel_scale; o[10] = cmd[1] * c->lin_vel_scale; o[11] = cmd[2] * c->ang_vel_scale;
        float tq2 = 0.0f;
        for (int j = 0; j < nd; j++) {
            o[12 + j] = (d[2 * j] - c->default_dof_pos[j]) * c->dof_pos_scale;
            o[12 + nd + j] = d[2 * j + 1] * c->dof_vel_scale;
            o[12 + 2 * nd + j] = act[j];
            const float t = torques[(size_t)e * nd + j];
            tq2 += t * t;
        }
        float* oc = obs_clamped + (size_t)e * no;
        for (int k = 0; k < no; k++) oc[k] = o[k] < -c->clip_obs ? -c->clip_obs : (o[k] > c->clip_obs ? c->clip_obs : o[k]);
        const float ex = cmd[0] - lin[0], ey = cmd[1] - lin[1], ez = cmd[2] - ang[2];
        float rw = expf(-(ex * ex + ey * ey) / 0.25f) * c->rew_lin_vel_xy + expf(-(ez * ez) / 0.25f) * c->rew_ang_vel_z + tq2 * c->rew_torque;
        rew[e] = rw > 0.0f ? rw : 0.0f;
        const float* cf = contact + (size_t)e * nb * 3;
        const float* b = cf + c->base_body * 3;
        int term = sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]) > 1.0f;
        for (int k = 0; k < c->n_knee; k++) {
            const float* g = cf + c->knee_bodies[k] * 3;
            term = term || (sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]) > 1.0f);
        }
        const int time_out = progress[e] >= c->max_episode_length - 1;
        reset[e] = (term || time_out) ? 1 : 0;
        timeout[e] = (time_out && reset[e]) ? 1 : 0;
    }
    return rc;
}
