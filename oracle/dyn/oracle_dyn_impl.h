/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
 * (isaacgymenv_b200/ + libb200gym.so).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, as the checker or the timed CPU baseline.
 *
 * CPU restatement (scalar C, float64 and float32 instantiations) of the physics that the reference
 * delegates to Isaac Gym / PhysX behind `gym.simulate` (reference call sites:
 * isaacgymenvs/tasks/base/vec_task.py:379-382, tasks/anymal_terrain.py:443-451; solver settings
 * cfg/task/Anymal.yaml:81-100; drives tasks/anymal.py:199-203).
 *
 * PARITY UNPINNED for dynamics: Isaac Gym Preview 4 (closed binary, un-vendored, un-pinned in
 * setup.py:15-25) is absent, and the reference holds no test or golden vector that touches the
 * simulator.  This file restates the *published* algorithms instead:
 *   - Featherstone articulated-body algorithm (forward dynamics), composite-rigid-body algorithm
 *     (mass matrix) and recursive Newton-Euler (bias forces), floating or fixed base
 *     [R. Featherstone, Rigid Body Dynamics Algorithms, 2008, ch. 5-7, 9];
 *   - implicit PD position drives folded into the joint-space inertia (h*Kd + h^2*Kp on the diagonal),
 *     which is how an implicit spring/damper drive enters a velocity-level articulation solver;
 *   - hard contact at the velocity level: projected relaxation over contact points with Coulomb friction (cone
 *     projection) -- Gauss-Seidel along a chain, Jacobi across chains (they couple only through the root) --
 *     impulses propagated through the articulated inertias, position
 *     iterations with penetration bias followed by velocity iterations without it (PGS/TGS-style
 *     split named by `num_position_iterations` / `num_velocity_iterations`);
 *   - semi-implicit Euler integration per sub-step.
 * It is self-validated by invariants in tests/ (ABA == CRBA/RNEA solve, kinetic energy against an
 * independent finite-difference FK, momentum and energy conservation), not by PhysX output, and pinned
 * against rigid-body mechanics derived independently of this file: the Euler-Lagrange equations of the
 * same robots by automatic differentiation of a kinematics-only Lagrangian (accelerations, mass matrix,
 * implicit-drive step, RK4 horizon: tests/test_oracle_lagrange.py) and the operational-space solution
 * W = J M^-1 J^T of one- and two-contact cases on the plane and on a sloped heightfield
 * (tests/test_oracle_contact_impulse.py).  What PhysX adds to mechanics stays unpinned.
 *
 * This header is a template: include it with ORC_REAL and ORC_SUF defined.
 */

#define ORC_CAT2(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT2(a, b)
#define FN(name) ORC_CAT(name, ORC_SUF)
#define R ORC_REAL

typedef struct FN(orc_kin) {
    R rot[B2G_MAX_LINKS][9];      /* link rotation, world axes                     */
    R pos[B2G_MAX_LINKS][3];      /* link origin relative to the root origin       */
    R S[B2G_MAX_DOF][6];          /* joint motion subspace (angular, linear at O)  */
    R vel[B2G_MAX_LINKS][6];      /* link spatial velocity                          */
    R cb[B2G_MAX_DOF][6];         /* velocity-product acceleration v x (S qd)       */
    R I[B2G_MAX_LINKS][36];       /* link spatial inertia about O                   */
    R IA[B2G_MAX_LINKS][36];      /* articulated inertia                            */
    R pA[B2G_MAX_LINKS][6];       /* articulated bias force                         */
    R U[B2G_MAX_DOF][6];
    R D[B2G_MAX_DOF];
    R u[B2G_MAX_DOF];
    R IA0inv[36];
} FN(orc_kin);

static void FN(cross3)(const R* a, const R* b, R* o) {
    R x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static void FN(matvec3)(const R* m, const R* v, R* o) {
    R x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
    R y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
    R z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void FN(matmul3)(const R* a, const R* b, R* o) {
    R t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
    for (int i = 0; i < 9; i++) o[i] = t[i];
}
static void FN(quat2mat)(const R* q, R* m) {
    R x = q[0], y = q[1], z = q[2], w = q[3];
    m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - z * w); m[2] = 2 * (x * z + y * w);
    m[3] = 2 * (x * y + z * w); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - x * w);
    m[6] = 2 * (x * z - y * w); m[7] = 2 * (y * z + x * w); m[8] = 1 - 2 * (x * x + y * y);
}
static void FN(axisangle2mat)(const R* a, R ang, R* m) {
    R c = (R)cos((double)ang), s = (R)sin((double)ang), t = 1 - c;
    m[0] = t * a[0] * a[0] + c; m[1] = t * a[0] * a[1] - s * a[2]; m[2] = t * a[0] * a[2] + s * a[1];
    m[3] = t * a[0] * a[1] + s * a[2]; m[4] = t * a[1] * a[1] + c; m[5] = t * a[1] * a[2] - s * a[0];
    m[6] = t * a[0] * a[2] - s * a[1]; m[7] = t * a[1] * a[2] + s * a[0]; m[8] = t * a[2] * a[2] + c;
}
/* motion cross product v x m */
static void FN(crm)(const R* v, const R* m, R* o) {
    R a[3], b[3], c[3];
    FN(cross3)(v, m, a); FN(cross3)(v, m + 3, b); FN(cross3)(v + 3, m, c);
    o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
/* force cross product v x* f */
static void FN(crf)(const R* v, const R* f, R* o) {
    R a[3], b[3], c[3];
    FN(cross3)(v, f, a); FN(cross3)(v + 3, f + 3, b); FN(cross3)(v, f + 3, c);
    o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
static void FN(mv6)(const R* m, const R* v, R* o) {
    R t[6];
    for (int i = 0; i < 6; i++) { R s = 0; for (int j = 0; j < 6; j++) s += m[i * 6 + j] * v[j]; t[i] = s; }
    for (int i = 0; i < 6; i++) o[i] = t[i];
}
static R FN(dot6)(const R* a, const R* b) { R s = 0; for (int i = 0; i < 6; i++) s += a[i] * b[i]; return s; }

/* spatial inertia about O from mass, com position c (rel. O) and rotational inertia about the com (world axes) */
static void FN(spatial_inertia)(R m, const R* c, const R* Ic, R* I) {
    R cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
    R cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            I[i * 6 + j] = Ic[i * 3 + j] + m * ((i == j ? cc : 0) - c[i] * c[j]);
            I[i * 6 + 3 + j] = m * cx[i * 3 + j];
            I[(3 + i) * 6 + j] = m * cx[j * 3 + i];
            I[(3 + i) * 6 + 3 + j] = (i == j) ? m : 0;
        }
}

/* inverse of a symmetric positive definite 6x6 via Cholesky; returns 0 on success */
static int FN(spd_inverse6)(const R* a, R* inv) {
    R L[36];
    for (int i = 0; i < 36; i++) L[i] = 0;
    for (int j = 0; j < 6; j++) {
        R d = a[j * 6 + j];
        for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k];
        if (!(d > 0)) return -1;
        R s = (R)sqrt((double)d);
        L[j * 6 + j] = s;
        for (int i = j + 1; i < 6; i++) {
            R t = a[i * 6 + j];
            for (int k = 0; k < j; k++) t -= L[i * 6 + k] * L[j * 6 + k];
            L[i * 6 + j] = t / s;
        }
    }
    for (int c = 0; c < 6; c++) {
        R y[6], x[6];
        for (int i = 0; i < 6; i++) {
            R t = (i == c) ? 1 : 0;
            for (int k = 0; k < i; k++) t -= L[i * 6 + k] * y[k];
            y[i] = t / L[i * 6 + i];
        }
        for (int i = 5; i >= 0; i--) {
            R t = y[i];
            for (int k = i + 1; k < 6; k++) t -= L[k * 6 + i] * x[k];
            x[i] = t / L[i * 6 + i];
        }
        for (int i = 0; i < 6; i++) inv[i * 6 + c] = x[i];
    }
    return 0;
}

/* ---- kinematics + velocity-dependent terms, common frame = world axes at the root origin ---- */
static void FN(orc_kinematics)(const b2g_model* m, const R* root13, const R* q, const R* qd, FN(orc_kin)* k) {
    int nd = m->n_dof;
    FN(quat2mat)(root13 + 3, k->rot[0]);
    k->pos[0][0] = k->pos[0][1] = k->pos[0][2] = 0;
    if (m->fixed_base) {
        for (int i = 0; i < 6; i++) k->vel[0][i] = 0;
    } else {
        for (int i = 0; i < 3; i++) { k->vel[0][i] = root13[10 + i]; k->vel[0][3 + i] = root13[7 + i]; }
    }
    for (int c = 0; c < m->n_chains; c++) {
        for (int j = 0; j < m->chain_len[c]; j++) {
            int d = m->chain_start[c] + j;
            int l = d + 1, p = (j == 0) ? 0 : l - 1;
            R jq[4] = {m->joint_quat[d][0], m->joint_quat[d][1], m->joint_quat[d][2], m->joint_quat[d][3]};
            R jp[3] = {m->joint_pos[d][0], m->joint_pos[d][1], m->joint_pos[d][2]};
            R ax[3] = {m->joint_axis[d][0], m->joint_axis[d][1], m->joint_axis[d][2]};
            R rj[9], rjw[9], off[3], axw[3];
            FN(quat2mat)(jq, rj);
            FN(matmul3)(k->rot[p], rj, rjw);
            FN(matvec3)(k->rot[p], jp, off);
            FN(matvec3)(rjw, ax, axw);
            R pj[3] = {k->pos[p][0] + off[0], k->pos[p][1] + off[1], k->pos[p][2] + off[2]};
            if (m->joint_type[d] == B2G_JOINT_REVOLUTE) {
                R rq[9];
                FN(axisangle2mat)(ax, q[d], rq);
                FN(matmul3)(rjw, rq, k->rot[l]);
                for (int i = 0; i < 3; i++) k->pos[l][i] = pj[i];
                R lin[3];
                FN(cross3)(pj, axw, lin);
                for (int i = 0; i < 3; i++) { k->S[d][i] = axw[i]; k->S[d][3 + i] = lin[i]; }
            } else {
                for (int i = 0; i < 9; i++) k->rot[l][i] = rjw[i];
                for (int i = 0; i < 3; i++) { k->pos[l][i] = pj[i] + axw[i] * q[d]; k->S[d][i] = 0; k->S[d][3 + i] = axw[i]; }
            }
            R vj[6];
            for (int i = 0; i < 6; i++) { vj[i] = k->S[d][i] * qd[d]; k->vel[l][i] = k->vel[p][i] + vj[i]; }
            FN(crm)(k->vel[l], vj, k->cb[d]);
        }
    }
    for (int l = 0; l <= nd; l++) {
        R com[3] = {m->link_com[l][0], m->link_com[l][1], m->link_com[l][2]};
        R il[9] = {m->link_inertia[l][0], m->link_inertia[l][3], m->link_inertia[l][4],
                   m->link_inertia[l][3], m->link_inertia[l][1], m->link_inertia[l][5],
                   m->link_inertia[l][4], m->link_inertia[l][5], m->link_inertia[l][2]};
        R cw[3], t[9], rt[9], iw[9];
        FN(matvec3)(k->rot[l], com, cw);
        for (int i = 0; i < 3; i++) cw[i] += k->pos[l][i];
        FN(matmul3)(k->rot[l], il, t);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rt[i * 3 + j] = k->rot[l][j * 3 + i];
        FN(matmul3)(t, rt, iw);
        FN(spatial_inertia)((R)m->link_mass[l], cw, iw, k->I[l]);
    }
}

/* ABA passes 1-2: articulated inertias and bias forces. tau = applied joint efforts, dext = extra
 * joint-space diagonal (armature + implicit drive terms). */
static int FN(orc_aba_backward)(const b2g_model* m, FN(orc_kin)* k, const R* tau, const R* dext) {
    int nd = m->n_dof;
    for (int l = 0; l <= nd; l++) {
        R iv[6];
        for (int i = 0; i < 36; i++) k->IA[l][i] = k->I[l][i];
        FN(mv6)(k->I[l], k->vel[l], iv);
        FN(crf)(k->vel[l], iv, k->pA[l]);
    }
    for (int c = 0; c < m->n_chains; c++) {
        for (int j = m->chain_len[c] - 1; j >= 0; j--) {
            int d = m->chain_start[c] + j;
            int l = d + 1, p = (j == 0) ? 0 : l - 1;
            FN(mv6)(k->IA[l], k->S[d], k->U[d]);
            k->D[d] = FN(dot6)(k->S[d], k->U[d]) + dext[d];
            k->u[d] = tau[d] - FN(dot6)(k->S[d], k->pA[l]);
            R Ia[36], pa[6], t[6];
            R dinv = 1 / k->D[d];
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) Ia[a * 6 + b] = k->IA[l][a * 6 + b] - k->U[d][a] * k->U[d][b] * dinv;
            FN(mv6)(Ia, k->cb[d], t);
            for (int a = 0; a < 6; a++) pa[a] = k->pA[l][a] + t[a] + k->U[d][a] * (k->u[d] * dinv);
            if (p != 0 || !m->fixed_base) {
                for (int a = 0; a < 36; a++) k->IA[p][a] += Ia[a];
                for (int a = 0; a < 6; a++) k->pA[p][a] += pa[a];
            }
        }
    }
    if (!m->fixed_base) return FN(spd_inverse6)(k->IA[0], k->IA0inv);
    for (int i = 0; i < 36; i++) k->IA0inv[i] = 0;
    return 0;
}

/* ABA pass 3. grav = gravity vector. Outputs qdd (nd) and the root's spatial acceleration (6). */
static void FN(orc_aba_forward)(const b2g_model* m, const FN(orc_kin)* k, const R* grav, R* qdd, R* a0) {
    R acc[B2G_MAX_LINKS][6];
    if (m->fixed_base) {
        for (int i = 0; i < 3; i++) { acc[0][i] = 0; acc[0][3 + i] = -grav[i]; }
    } else {
        /* gravity is a uniform field: solve for a' = a - a_g (pA holds no gravity term), add a_g back below */
        FN(mv6)(k->IA0inv, k->pA[0], acc[0]);
        for (int i = 0; i < 6; i++) acc[0][i] = -acc[0][i];
    }
    for (int c = 0; c < m->n_chains; c++) {
        for (int j = 0; j < m->chain_len[c]; j++) {
            int d = m->chain_start[c] + j;
            int l = d + 1, p = (j == 0) ? 0 : l - 1;
            R ap[6];
            for (int i = 0; i < 6; i++) ap[i] = acc[p][i] + k->cb[d][i];
            qdd[d] = (k->u[d] - FN(dot6)(k->U[d], ap)) / k->D[d];
            for (int i = 0; i < 6; i++) acc[l][i] = ap[i] + k->S[d][i] * qdd[d];
        }
    }
    /* true root acceleration = relative acceleration + gravity field (zero for a fixed base) */
    for (int i = 0; i < 3; i++) { a0[i] = acc[0][i]; a0[3 + i] = acc[0][3 + i] + grav[i]; }
}

/* velocity change caused by a spatial impulse F applied to link `link` (0 = root) plus optional joint
 * impulse tj on DOF dj (dj < 0: none). Accumulates into dv0 (6) and dqd (nd). Uses IA/U/D from the
 * last orc_aba_backward. */
static void FN(orc_apply_impulse)(const b2g_model* m, const FN(orc_kin)* k, int link, const R* F, int dj, R tj, R* v0, R* qd) {
    R ud[B2G_MAX_DOF];
    R P[6] = {0, 0, 0, 0, 0, 0};
    for (int d = 0; d < m->n_dof; d++) ud[d] = 0;
    int cpath = -1;
    if (link > 0 || dj >= 0) {
        int dd = (link > 0) ? link - 1 : dj;
        for (int c = 0; c < m->n_chains; c++)
            if (dd >= m->chain_start[c] && dd < m->chain_start[c] + m->chain_len[c]) cpath = c;
    }
    if (link == 0 && F) for (int i = 0; i < 6; i++) P[i] = -F[i];
    if (cpath >= 0) {
        int c = cpath;
        int have = 0;
        for (int j = m->chain_len[c] - 1; j >= 0; j--) {
            int d = m->chain_start[c] + j;
            int l = d + 1;
            if (F && l == link) { for (int i = 0; i < 6; i++) P[i] -= F[i]; have = 1; }
            if (d == dj) have = 1;
            if (!have) continue;
            ud[d] = ((d == dj) ? tj : 0) - FN(dot6)(k->S[d], P);
            R s = ud[d] / k->D[d];
            for (int i = 0; i < 6; i++) P[i] += k->U[d][i] * s;
        }
    }
    R dv[B2G_MAX_LINKS][6];
    if (m->fixed_base) {
        for (int i = 0; i < 6; i++) dv[0][i] = 0;
    } else {
        FN(mv6)(k->IA0inv, P, dv[0]);
        for (int i = 0; i < 6; i++) { dv[0][i] = -dv[0][i]; v0[i] += dv[0][i]; }
    }
    for (int c = 0; c < m->n_chains; c++) {
        for (int j = 0; j < m->chain_len[c]; j++) {
            int d = m->chain_start[c] + j;
            int l = d + 1, p = (j == 0) ? 0 : l - 1;
            R dq = (ud[d] - FN(dot6)(k->U[d], dv[p])) / k->D[d];
            for (int i = 0; i < 6; i++) dv[l][i] = dv[p][i] + k->S[d][i] * dq;
            qd[d] += dq;
        }
    }
}

/* spatial velocity of a link from (v0, qd) */
static void FN(orc_link_velocity)(const b2g_model* m, const FN(orc_kin)* k, int link, const R* v0, const R* qd, R* v) {
    for (int i = 0; i < 6; i++) v[i] = m->fixed_base ? 0 : v0[i];
    if (link == 0) return;
    int dd = link - 1;
    for (int c = 0; c < m->n_chains; c++) {
        if (dd < m->chain_start[c] || dd >= m->chain_start[c] + m->chain_len[c]) continue;
        for (int d = m->chain_start[c]; d <= dd; d++)
            for (int i = 0; i < 6; i++) v[i] += k->S[d][i] * qd[d];
    }
}

/* terrain height and unit normal under world point (x, y); plane z = 0 when there is no heightfield */
static void FN(orc_ground)(const b2g_heightfield* hf, const int16_t* s, R x, R y, R* h, R* n) {
    n[0] = 0; n[1] = 0; n[2] = 1; *h = 0;
    if (!hf || !s) return;
    R gx = (x - (R)hf->origin_x) / (R)hf->horizontal_scale, gy = (y - (R)hf->origin_y) / (R)hf->horizontal_scale;
    if (gx < 0) gx = 0; if (gy < 0) gy = 0;
    if (gx > (R)(hf->rows - 1)) gx = (R)(hf->rows - 1);
    if (gy > (R)(hf->cols - 1)) gy = (R)(hf->cols - 1);
    int i = (int)gx, j = (int)gy;
    if (i > hf->rows - 2) i = hf->rows - 2;
    if (j > hf->cols - 2) j = hf->cols - 2;
    R fx = gx - (R)i, fy = gy - (R)j;
    R vs = (R)hf->vertical_scale, hs = (R)hf->horizontal_scale;
    R h00 = vs * (R)s[i * hf->cols + j], h10 = vs * (R)s[(i + 1) * hf->cols + j];
    R h01 = vs * (R)s[i * hf->cols + j + 1], h11 = vs * (R)s[(i + 1) * hf->cols + j + 1];
    R dzdx, dzdy;
    if (fx >= fy) { dzdx = (h10 - h00); dzdy = (h11 - h10); *h = h00 + dzdx * fx + dzdy * fy; }
    else { dzdx = (h11 - h01); dzdy = (h01 - h00); *h = h00 + dzdx * fx + dzdy * fy; }
    R nx = -dzdx / hs, ny = -dzdy / hs, nz = 1;
    R inv = 1 / (R)sqrt((double)(nx * nx + ny * ny + nz * nz));
    n[0] = nx * inv; n[1] = ny * inv; n[2] = nz * inv;
}

typedef struct FN(orc_contact) {
    int link, body;
    int self_c;      /* 1: the other body is the robot's own root link (self-collision), 0: the ground */
    R r[3];          /* contact point relative to O */
    R n[3], t1[3], t2[3];
    R gap;
    R A[9];          /* local Delassus block in (n, t1, t2) */
    R lam[3];        /* accumulated impulse (n, t1, t2) */
} FN(orc_contact);

static void FN(orc_contact_wrench)(const FN(orc_contact)* c, const R* dir, R* F) {
    R mom[3];
    FN(cross3)(c->r, dir, mom);
    for (int i = 0; i < 3; i++) { F[i] = mom[i]; F[3 + i] = dir[i]; }
}

static void FN(orc_point_velocity)(const FN(orc_contact)* c, const R* v, R* out);
static void FN(orc_contact_wrench)(const FN(orc_contact)* c, const R* dir, R* F);

/* impulse F on the contact's link; a self contact puts the reaction -F on the root */
static void FN(orc_apply_contact_impulse)(const b2g_model* m, const FN(orc_kin)* k, const FN(orc_contact)* cc, const R* F, R* v0, R* qd) {
    FN(orc_apply_impulse)(m, k, cc->link, F, -1, 0, v0, qd);
    if (cc->self_c && !m->fixed_base) {
        R Fm[6];
        for (int i = 0; i < 6; i++) Fm[i] = -F[i];
        FN(orc_apply_impulse)(m, k, 0, Fm, -1, 0, v0, qd);
    }
}

/* velocity of the contact point on the link relative to the other body: the ground (at rest) or, for a self contact, the root */
static void FN(orc_contact_velocity)(const b2g_model* m, const FN(orc_kin)* k, const FN(orc_contact)* cc, const R* v0, const R* qd, R* pv) {
    R lv[6];
    FN(orc_link_velocity)(m, k, cc->link, v0, qd, lv);
    if (cc->self_c && !m->fixed_base) for (int i = 0; i < 6; i++) lv[i] -= v0[i];
    FN(orc_point_velocity)(cc, lv, pv);
}

/* tangent frame, zero impulse and local Delassus block of a contact whose link, point, normal and gap are set */
static void FN(orc_finish_contact)(const b2g_model* m, const FN(orc_kin)* k, FN(orc_contact)* cc) {
    int nd = m->n_dof;
    for (int a = 0; a < 3; a++) cc->lam[a] = 0;
    /* tangent basis: t1 = normalise(x_world - (x.n) n), t2 = n x t1 */
    R dn = cc->n[0];
    R t1[3] = {1 - dn * cc->n[0], -dn * cc->n[1], -dn * cc->n[2]};
    R inv = 1 / (R)sqrt((double)(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]));
    for (int a = 0; a < 3; a++) cc->t1[a] = t1[a] * inv;
    FN(cross3)(cc->n, cc->t1, cc->t2);
    const R* dirs[3] = {cc->n, cc->t1, cc->t2};
    for (int b = 0; b < 3; b++) {
        R F[6], dv0[6] = {0, 0, 0, 0, 0, 0}, dqd[B2G_MAX_DOF], pv[3];
        for (int d = 0; d < nd; d++) dqd[d] = 0;
        FN(orc_contact_wrench)(cc, dirs[b], F);
        FN(orc_apply_contact_impulse)(m, k, cc, F, dv0, dqd);
        FN(orc_contact_velocity)(m, k, cc, dv0, dqd, pv);
        for (int a = 0; a < 3; a++) cc->A[a * 3 + b] = pv[0] * dirs[a][0] + pv[1] * dirs[a][1] + pv[2] * dirs[a][2];
    }
}

/* Self-collision (b2g_sim_params::self_collision; the reference enables it for the rough-terrain tasks, tasks/anymal_terrain.py:282):
 * the candidate spheres of every link that is not attached to the root directly (PhysX does not collide a link with its parent)
 * against the BASE's bounding box -- the axis-aligned box, in the root frame, around the candidates of API body 0 and their radii.
 * Signed distance = the largest of the three slab distances (exact over a face, a lower bound next to an edge), normal = that face's
 * outward normal, contact point on the sphere's surface.  Returns 1 and fills link / point / normal / gap when inside the offset. */
static void FN(orc_root_box)(const b2g_model* m, R* c, R* hx) {
    R lo[3] = {(R)1e30, (R)1e30, (R)1e30}, hi[3] = {(R)-1e30, (R)-1e30, (R)-1e30};
    int cnt = 0;
    for (int i = 0; i < m->n_cpts; i++) {
        if (m->cp_link[i] != 0 || m->cp_body[i] != 0) continue;      /* the base body itself, not the bodies fixed to it */
        cnt++;
        for (int a = 0; a < 3; a++) {
            R p = (R)m->cp_pos[i][a], r = (R)m->cp_radius[i];
            if (p - r < lo[a]) lo[a] = p - r;
            if (p + r > hi[a]) hi[a] = p + r;
        }
    }
    for (int a = 0; a < 3; a++) { c[a] = cnt ? (R)0.5 * (lo[a] + hi[a]) : 0; hx[a] = cnt ? (R)0.5 * (hi[a] - lo[a]) : (R)-1e30; }
}
static int FN(orc_self_candidate)(const b2g_model* m, const FN(orc_kin)* k, int i, R offset, const R* bc, const R* bh, FN(orc_contact)* cc) {
    int l = m->cp_link[i];
    R lp[3] = {m->cp_pos[i][0], m->cp_pos[i][1], m->cp_pos[i][2]}, rc[3], q[3];
    FN(matvec3)(k->rot[l], lp, rc);
    for (int a = 0; a < 3; a++) rc[a] += k->pos[l][a];
    for (int a = 0; a < 3; a++) q[a] = k->rot[0][0 * 3 + a] * rc[0] + k->rot[0][1 * 3 + a] * rc[1] + k->rot[0][2 * 3 + a] * rc[2] - bc[a];   /* R0^T rc */
    R rad = (R)m->cp_radius[i];
    int best = 0;
    R gap = (q[0] < 0 ? -q[0] : q[0]) - bh[0] - rad;
    for (int a = 1; a < 3; a++) {
        R d = (q[a] < 0 ? -q[a] : q[a]) - bh[a] - rad;
        if (d > gap) { gap = d; best = a; }
    }
    if (gap >= offset) return 0;
    R sgn = q[best] < 0 ? (R)-1 : (R)1;
    cc->link = l; cc->body = m->cp_body[i]; cc->gap = gap; cc->self_c = 1;
    for (int a = 0; a < 3; a++) { cc->n[a] = sgn * k->rot[0][a * 3 + best]; cc->r[a] = rc[a] - rad * cc->n[a]; }
    return 1;
}

static void FN(orc_point_velocity)(const FN(orc_contact)* c, const R* v, R* out) {
    R wxr[3];
    FN(cross3)(v, c->r, wxr);
    for (int i = 0; i < 3; i++) out[i] = v[3 + i] + wxr[i];
}

/*
 * One sub-step of length h for one environment.
 *  root13: pos3 quat4 linvel3 angvel3 (in/out)   dof: nd x (pos, vel) (in/out)
 *  target: position (POS) / velocity (VEL) targets; actuation: efforts (EFFORT mode)
 *  dof_force (nd) and contact (nb x 3, world) are outputs.
 */
static int FN(orc_substep)(const b2g_model* m, const b2g_sim_params* sp, const b2g_dof_props* dp,
                           const b2g_heightfield* hf, const int16_t* hfs, R mu_shape, R h,
                           R* root13, R* dof, const R* target, const R* actuation, R* dof_force, R* contact) {
    int nd = m->n_dof;
    FN(orc_kin)* k = (FN(orc_kin)*)malloc(sizeof(FN(orc_kin)));
    if (!k) return -1;
    R q[B2G_MAX_DOF], qd[B2G_MAX_DOF], tau[B2G_MAX_DOF], dext[B2G_MAX_DOF], qdd[B2G_MAX_DOF], a0[6];
    R grav[3] = {(R)sp->gravity[0], (R)sp->gravity[1], (R)sp->gravity[2]};
    for (int d = 0; d < nd; d++) { q[d] = dof[2 * d]; qd[d] = dof[2 * d + 1]; }
    for (int d = 0; d < nd; d++) {
        R kp = (R)dp->stiffness[d], kd = (R)dp->damping[d];
        dext[d] = (R)m->armature[d];
        tau[d] = 0;
        /* implicit PD / velocity drive; a drive whose torque -- estimated at the end-of-step position with the current velocity -- exceeds
         * the effort limit is saturated: a constant torque of that size, no implicit terms (PhysX clamps the drive force to maxForce) */
        R lim_d = (R)dp->effort[d];
        if (dp->drive_mode[d] == B2G_DOF_MODE_POS) {
            R te = kp * (target[d] - q[d] - h * qd[d]) - kd * qd[d];
            if (lim_d > 0 && (te > lim_d || te < -lim_d)) {
                tau[d] = te > 0 ? lim_d : -lim_d;
            } else {
                tau[d] = kp * (target[d] - q[d]) - (kd + h * kp) * qd[d];
                dext[d] += h * kd + h * h * kp;
            }
        } else if (dp->drive_mode[d] == B2G_DOF_MODE_VEL) {
            R te = kd * (target[d] - qd[d]);
            if (lim_d > 0 && (te > lim_d || te < -lim_d)) {
                tau[d] = te > 0 ? lim_d : -lim_d;
            } else {
                tau[d] = te;
                dext[d] += h * kd;
            }
        } else if (dp->drive_mode[d] == B2G_DOF_MODE_EFFORT) {
            R e = actuation[d], lim = (R)dp->effort[d];
            if (lim > 0) { if (e > lim) e = lim; if (e < -lim) e = -lim; }
            tau[d] = e;
        }
        /* joint limits: one-sided implicit spring-damper (same implicit folding as the drive), active when the joint
         * is beyond a limit or would cross it within this sub-step at its current velocity */
        {
            R lo = (R)dp->lower[d], hi = (R)dp->upper[d], kl = (R)sp->joint_limit_stiffness, dl = (R)sp->joint_limit_damping;
            R qp = q[d] + h * qd[d];
            R ref = 0; int on = 0;
            if (lo > -1e30f && (q[d] < lo || qp < lo)) { ref = lo; on = 1; }
            else if (hi < 1e30f && (q[d] > hi || qp > hi)) { ref = hi; on = 1; }
            if (on) {
                tau[d] += kl * (ref - q[d]) - (dl + h * kl) * qd[d];
                dext[d] += h * dl + h * h * kl;
            }
        }
    }
    FN(orc_kinematics)(m, root13, q, qd, k);
    if (FN(orc_aba_backward)(m, k, tau, dext) != 0) { free(k); return -2; }
    FN(orc_aba_forward)(m, k, grav, qdd, a0);

    /* free velocity */
    R v0[6] = {0, 0, 0, 0, 0, 0};
    if (!m->fixed_base) {
        R w[3] = {k->vel[0][0], k->vel[0][1], k->vel[0][2]}, vl[3] = {k->vel[0][3], k->vel[0][4], k->vel[0][5]}, wxv[3];
        FN(cross3)(w, vl, wxv);
        /* root origin is a body-fixed point: classical acceleration = spatial + w x v */
        for (int i = 0; i < 3; i++) { v0[i] = w[i] + h * a0[i]; v0[3 + i] = vl[i] + h * (a0[3 + i] + wxv[i]); }
    }
    for (int d = 0; d < nd; d++) qd[d] += h * qdd[d];

    /* ---- collect contacts: per chain, candidates in model order (tip first), the first max_contacts_per_chain of them.
     * With B2G_SEGMENTS=1 in the environment, floating-base robots whose chains, cut into pieces of at most three links, fit
     * eight lanes run the segment kernels (b2g_host_pack.h::build_segments): there every PIECE has max_contacts_per_chain
     * slots of its own (root candidates count with the proximal piece); the order of a chain's contacts -- distal piece
     * first -- is the same. ---- */
    FN(orc_contact) con[B2G_MAX_CHAINS][2 * B2G_MAX_CONTACTS_PER_CHAIN];
    int ncon[B2G_MAX_CHAINS], npiece[B2G_MAX_CHAINS][2], cpiece[B2G_MAX_CHAINS][2 * B2G_MAX_CONTACTS_PER_CHAIN];
    int pieces = m->n_chains;
    for (int c = 0; c < m->n_chains; c++) pieces += m->chain_len[c] > 3 ? 1 : 0;
    const char* seg_env = getenv("B2G_SEGMENTS");
    const int by_piece = seg_env && seg_env[0] == '1' && !m->fixed_base && m->n_chains > 0 && pieces <= B2G_MAX_CHAINS;
    R mu_g = hf && hfs ? (R)hf->friction : (R)sp->plane_dynamic_friction;
    R mu = (R)0.5 * (mu_g + mu_shape);   /* PhysX default combine mode: average */
    int ground = (hf && hfs) || sp->has_ground;
    int maxc = sp->max_contacts_per_chain > 0 ? sp->max_contacts_per_chain : B2G_DEFAULT_CONTACTS_PER_CHAIN;
    if (maxc > B2G_MAX_CONTACTS_PER_CHAIN) maxc = B2G_MAX_CONTACTS_PER_CHAIN;
    for (int c = 0; c < m->n_chains; c++) { ncon[c] = 0; npiece[c][0] = npiece[c][1] = 0; }
    R box_c[3], box_h[3];
    FN(orc_root_box)(m, box_c, box_h);
    /* candidates come grouped by link (tip first within a chain, root candidates last): per link the ground tests, then -- links that do
     * not hang off the root directly -- the self-collision tests against the root's box; both kinds share the chain's (piece's) slots */
    for (int i0 = 0; i0 < m->n_cpts;) {
        int c = m->cp_chain[i0], l = m->cp_link[i0], i1 = i0;
        while (i1 < m->n_cpts && m->cp_link[i1] == l && m->cp_chain[i1] == c) i1++;
        const int piece = (by_piece && l > 0 && l - 1 - m->chain_start[c] >= 3) ? 1 : 0;
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 0 && (!ground || (m->fixed_base && l == 0))) continue;
            if (pass == 1 && !(sp->self_collision && l > 0 && l - 1 - m->chain_start[c] >= 1)) continue;
            for (int i = i0; i < i1; i++) {
                /* slots full: the candidate goes into a scratch record first and replaces the chain's (piece's) shallowest contact only
                 * if it penetrates further -- the deepest candidates are the ones kept */
                const int full = by_piece ? npiece[c][piece] >= maxc : ncon[c] >= maxc;
                FN(orc_contact) spare;
                FN(orc_contact)* cc = full ? &spare : &con[c][ncon[c]];
                if (pass == 0) {
                    R lp[3] = {m->cp_pos[i][0], m->cp_pos[i][1], m->cp_pos[i][2]}, rc[3], gh, n[3];
                    FN(matvec3)(k->rot[l], lp, rc);
                    for (int a = 0; a < 3; a++) rc[a] += k->pos[l][a];
                    FN(orc_ground)(hf, hfs, root13[0] + rc[0], root13[1] + rc[1], &gh, n);
                    R gap = (root13[2] + rc[2] - gh) * n[2] - (R)m->cp_radius[i];
                    if (gap >= (R)sp->contact_offset) continue;
                    cc->link = l; cc->body = m->cp_body[i]; cc->gap = gap; cc->self_c = 0;
                    for (int a = 0; a < 3; a++) { cc->n[a] = n[a]; cc->r[a] = rc[a] - (R)m->cp_radius[i] * n[a]; }
                } else if (!FN(orc_self_candidate)(m, k, i, (R)sp->contact_offset, box_c, box_h, cc)) {
                    continue;
                }
                if (full) {
                    int w = -1;
                    for (int q = 0; q < ncon[c]; q++) {
                        if (by_piece && cpiece[c][q] != piece) continue;
                        if (w < 0 || con[c][q].gap > con[c][w].gap) w = q;
                    }
                    if (w < 0 || !(cc->gap < con[c][w].gap)) continue;
                    con[c][w] = *cc;
                    cc = &con[c][w];
                } else {
                    cpiece[c][ncon[c]] = piece;
                    ncon[c]++;
                    npiece[c][piece]++;
                }
                FN(orc_finish_contact)(m, k, cc);
            }
        }
        i0 = i1;
    }
    /* ---- projected Gauss-Seidel: position iterations (with bias), integrate, velocity iterations ---- */
    int npos = sp->num_position_iterations, nvel = sp->num_velocity_iterations;
    R maxdep = (R)sp->max_depenetration_velocity;
    R vpos0[6], qdpos[B2G_MAX_DOF];
    for (int it = 0; it <= npos + nvel; it++) {
        if (it == npos) {   /* snapshot the velocity used to integrate positions */
            for (int i = 0; i < 6; i++) vpos0[i] = v0[i];
            for (int d = 0; d < nd; d++) qdpos[d] = qd[d];
        }
        if (it == npos + nvel) break;
        int with_bias = it < npos;
        /* slot s of every chain is updated from the SAME velocities (Jacobi across chains: they couple only through
         * the root), then all impulses are applied; slots of one chain follow each other (Gauss-Seidel within a chain) */
        for (int s = 0; s < 2 * B2G_MAX_CONTACTS_PER_CHAIN; s++) {
            R Fall[B2G_MAX_CHAINS][6];
            for (int c = 0; c < m->n_chains; c++) {
                for (int a = 0; a < 6; a++) Fall[c][a] = 0;
                if (s >= ncon[c]) continue;
                FN(orc_contact)* cc = &con[c][s];
                R pv[3];
                FN(orc_contact_velocity)(m, k, cc, v0, qd, pv);
                R vn = pv[0] * cc->n[0] + pv[1] * cc->n[1] + pv[2] * cc->n[2];
                R vt1 = pv[0] * cc->t1[0] + pv[1] * cc->t1[1] + pv[2] * cc->t1[2];
                R vt2 = pv[0] * cc->t2[0] + pv[1] * cc->t2[1] + pv[2] * cc->t2[2];
                R tgt = -cc->gap / h;
                if (tgt > maxdep) tgt = maxdep;
                if (!with_bias && tgt > 0) tgt = 0;
                R dl[3];
                R ln = cc->lam[0] - (vn - tgt) / cc->A[0];
                if (ln < 0) ln = 0;
                dl[0] = ln - cc->lam[0];
                vt1 += cc->A[3] * dl[0]; vt2 += cc->A[6] * dl[0];
                R l1 = cc->lam[1] - vt1 / cc->A[4];
                R vt2s = vt2 + cc->A[7] * (l1 - cc->lam[1]);
                R l2 = cc->lam[2] - vt2s / cc->A[8];
                R lim_t = mu * ln, mag = (R)sqrt((double)(l1 * l1 + l2 * l2));
                if (mag > lim_t) {
                    /* the sticking impulse leaves the cone -> the contact slides.  Scaling the sticking impulse back would keep ITS
                     * direction (A_tt^-1 v_t), which is not opposite to the sliding velocity when the tangential Delassus block is
                     * anisotropic (a corner of a box, a foot on a leg).  A proximal step with a SCALAR step length followed by the
                     * radial projection has the Coulomb law as its fixed point: friction of magnitude mu * lambda_n opposite to the
                     * tangential velocity (maximum dissipation).  Checked by the block-on-a-slope known-answer test. */
                    R ia = 1 / ((cc->A[4] > cc->A[8] ? cc->A[4] : cc->A[8]) + (cc->A[7] < 0 ? -cc->A[7] : cc->A[7]));
                    l1 = cc->lam[1] - vt1 * ia;
                    l2 = cc->lam[2] - vt2 * ia;
                    mag = (R)sqrt((double)(l1 * l1 + l2 * l2));
                    if (mag > lim_t) { R sc = (mag > 0) ? lim_t / mag : 0; l1 *= sc; l2 *= sc; }
                }
                dl[1] = l1 - cc->lam[1]; dl[2] = l2 - cc->lam[2];
                cc->lam[0] = ln; cc->lam[1] = l1; cc->lam[2] = l2;
                R dir[3];
                for (int a = 0; a < 3; a++) dir[a] = cc->n[a] * dl[0] + cc->t1[a] * dl[1] + cc->t2[a] * dl[2];
                FN(orc_contact_wrench)(cc, dir, Fall[c]);
            }
            for (int c = 0; c < m->n_chains; c++) {
                if (s >= ncon[c]) continue;
                FN(orc_apply_contact_impulse)(m, k, &con[c][s], Fall[c], v0, qd);
            }
        }
    }
    if (npos + nvel == 0) {
        for (int i = 0; i < 6; i++) vpos0[i] = v0[i];
        for (int d = 0; d < nd; d++) qdpos[d] = qd[d];
    }

    if (!m->fixed_base) {   /* root velocity limits (b2g_sim_params::max_linear_velocity / max_angular_velocity; 0 = none) */
        R* vs[2] = {v0, vpos0};
        for (int k = 0; k < 2; k++) {
            R* v = vs[k];
            R lim_w = (R)sp->max_angular_velocity, lim_v = (R)sp->max_linear_velocity;
            R w2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], l2 = v[3] * v[3] + v[4] * v[4] + v[5] * v[5];
            if (lim_w > 0 && w2 > lim_w * lim_w) { R sc = lim_w / (R)sqrt((double)w2); v[0] *= sc; v[1] *= sc; v[2] *= sc; }
            if (lim_v > 0 && l2 > lim_v * lim_v) { R sc = lim_v / (R)sqrt((double)l2); v[3] *= sc; v[4] *= sc; v[5] *= sc; }
        }
    }
    /* ---- integrate positions with the post-position-iteration velocity ---- */
    for (int d = 0; d < nd; d++) {
        R vl = (R)dp->velocity[d];
        if (vl > 0) { if (qd[d] > vl) qd[d] = vl; if (qd[d] < -vl) qd[d] = -vl; if (qdpos[d] > vl) qdpos[d] = vl; if (qdpos[d] < -vl) qdpos[d] = -vl; }
        q[d] += h * qdpos[d];
        dof[2 * d] = q[d]; dof[2 * d + 1] = qd[d];
    }
    if (!m->fixed_base) {
        for (int i = 0; i < 3; i++) root13[i] += h * vpos0[3 + i];
        R w[3] = {vpos0[0], vpos0[1], vpos0[2]};
        R ang = (R)sqrt((double)(w[0] * w[0] + w[1] * w[1] + w[2] * w[2])) * h;
        R dq[4] = {0, 0, 0, 1};
        if (ang > (R)1e-12) {
            R s = (R)sin((double)(ang / 2)) / (ang / h);
            dq[0] = w[0] * s; dq[1] = w[1] * s; dq[2] = w[2] * s; dq[3] = (R)cos((double)(ang / 2));
        }
        R* qo = root13 + 3;
        R x1 = dq[0], y1 = dq[1], z1 = dq[2], w1 = dq[3], x2 = qo[0], y2 = qo[1], z2 = qo[2], w2 = qo[3];
        R nq[4] = {w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                   w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2};
        R nn = 1 / (R)sqrt((double)(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]));
        for (int i = 0; i < 4; i++) qo[i] = nq[i] * nn;
        for (int i = 0; i < 3; i++) { root13[7 + i] = v0[3 + i]; root13[10 + i] = v0[i]; }
    }
    /* ---- outputs ---- */
    for (int d = 0; d < nd; d++) {
        R f = 0, lim_e = (R)dp->effort[d];
        if (dp->drive_mode[d] == B2G_DOF_MODE_POS) f = (R)dp->stiffness[d] * (target[d] - q[d]) - (R)dp->damping[d] * qd[d];
        else if (dp->drive_mode[d] == B2G_DOF_MODE_VEL) f = (R)dp->damping[d] * (target[d] - qd[d]);
        else if (dp->drive_mode[d] == B2G_DOF_MODE_EFFORT) f = actuation[d];
        if (lim_e > 0) { if (f > lim_e) f = lim_e; if (f < -lim_e) f = -lim_e; }
        dof_force[d] = f;
    }
    for (int b = 0; b < m->n_bodies * 3; b++) contact[b] = 0;
    for (int c = 0; c < m->n_chains; c++)
        for (int s = 0; s < ncon[c]; s++) {
            FN(orc_contact)* cc = &con[c][s];
            for (int a = 0; a < 3; a++) {
                R f = (cc->n[a] * cc->lam[0] + cc->t1[a] * cc->lam[1] + cc->t2[a] * cc->lam[2]) / h;
                contact[cc->body * 3 + a] += f;
                if (cc->self_c) contact[0 * 3 + a] -= f;      /* the reaction on the root body (API body 0) */
            }
        }
    free(k);
    return 0;
}

/* gym.simulate for n_envs environments: sp->substeps sub-steps of dt/substeps each. */
int FN(orc_simulate)(const b2g_model* m, const b2g_sim_params* sp, const b2g_dof_props* dp,
                     const b2g_heightfield* hf, const int16_t* hfs, const float* friction, int n_envs,
                     R* root, R* dof, const R* target, const R* actuation, R* dof_force, R* contact) {
    int nd = m->n_dof, nb = m->n_bodies;
    R h = (R)sp->dt / (R)(sp->substeps > 0 ? sp->substeps : 1);
    int rc = 0;
    for (int e = 0; e < n_envs; e++) {
        R dummy_root[13] = {0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0};
        R* r = root ? root + (size_t)e * 13 : dummy_root;
        for (int s = 0; s < (sp->substeps > 0 ? sp->substeps : 1); s++) {
            int st = FN(orc_substep)(m, sp, dp, hf, hfs, friction ? (R)friction[e] : (R)1, h, r, dof + (size_t)e * nd * 2,
                                     target + (size_t)e * nd, actuation + (size_t)e * nd, dof_force + (size_t)e * nd,
                                     contact + (size_t)e * nb * 3);
            if (st != 0) rc = st;
        }
    }
    return rc;
}

/*
 * CONVERGED REFERENCE of one sub-step (what the production solver is measured against; tests/test_solver_convergence.py,
 * DESIGN.md section 6).  Same free dynamics, same contact model (spheres against the plane / heightfield, Coulomb cone, the same
 * penetration bias and position / velocity phases) -- but none of the production solver's shortcuts:
 *   - EVERY candidate inside the contact offset becomes a contact (no B2G_MAX_CONTACTS_PER_CHAIN cap);
 *   - plain sequential Gauss-Seidel over all contacts in one list (no Jacobi split across chains): each impulse is applied
 *     before the next contact reads its velocity;
 *   - each phase is iterated to convergence (max |delta lambda| <= tol (1 + max |lambda|), at most max_iter sweeps, >= 200 asked
 *     for by the tests) instead of num_position_iterations + num_velocity_iterations sweeps;
 *   - flags & 1: joint limits are HARD unilateral joint-space rows solved in the same sweeps (what PhysX does; the
 *     reference configures them through the URDF limits, assets/urdf/anymal_c/urdf/anymal.urdf:624-631) instead of the
 *     production path's one-sided implicit spring-dampers.
 * It is still this repo's contact model, not PhysX: it bounds the error of the SOLVER, not of the model.
 * info[0] = contacts, info[1] / info[2] = sweeps used by the position / velocity phase, info[3] = 1 if a phase hit max_iter.
 */
static int FN(orc_substep_ref)(const b2g_model* m, const b2g_sim_params* sp, const b2g_dof_props* dp,
                               const b2g_heightfield* hf, const int16_t* hfs, R mu_shape, R h,
                               R* root13, R* dof, const R* target, const R* actuation, R* dof_force, R* contact,
                               int flags, int max_iter, R tol, int* info) {
    int nd = m->n_dof;
    const int hard = flags & 1;
    FN(orc_kin)* k = (FN(orc_kin)*)malloc(sizeof(FN(orc_kin)));
    FN(orc_contact)* con = (FN(orc_contact)*)malloc(sizeof(FN(orc_contact)) * 2 * B2G_MAX_CPTS);
    if (!k || !con) { free(k); free(con); return -1; }
    R q[B2G_MAX_DOF], qd[B2G_MAX_DOF], tau[B2G_MAX_DOF], dext[B2G_MAX_DOF], qdd[B2G_MAX_DOF], a0[6];
    R grav[3] = {(R)sp->gravity[0], (R)sp->gravity[1], (R)sp->gravity[2]};
    for (int d = 0; d < nd; d++) { q[d] = dof[2 * d]; qd[d] = dof[2 * d + 1]; }
    for (int d = 0; d < nd; d++) {
        R kp = (R)dp->stiffness[d], kd = (R)dp->damping[d];
        dext[d] = (R)m->armature[d];
        tau[d] = 0;
        /* implicit PD / velocity drive; a drive whose torque -- estimated at the end-of-step position with the current velocity -- exceeds
         * the effort limit is saturated: a constant torque of that size, no implicit terms (PhysX clamps the drive force to maxForce) */
        R lim_d = (R)dp->effort[d];
        if (dp->drive_mode[d] == B2G_DOF_MODE_POS) {
            R te = kp * (target[d] - q[d] - h * qd[d]) - kd * qd[d];
            if (lim_d > 0 && (te > lim_d || te < -lim_d)) {
                tau[d] = te > 0 ? lim_d : -lim_d;
            } else {
                tau[d] = kp * (target[d] - q[d]) - (kd + h * kp) * qd[d];
                dext[d] += h * kd + h * h * kp;
            }
        } else if (dp->drive_mode[d] == B2G_DOF_MODE_VEL) {
            R te = kd * (target[d] - qd[d]);
            if (lim_d > 0 && (te > lim_d || te < -lim_d)) {
                tau[d] = te > 0 ? lim_d : -lim_d;
            } else {
                tau[d] = te;
                dext[d] += h * kd;
            }
        } else if (dp->drive_mode[d] == B2G_DOF_MODE_EFFORT) {
            R e = actuation[d], lim = (R)dp->effort[d];
            if (lim > 0) { if (e > lim) e = lim; if (e < -lim) e = -lim; }
            tau[d] = e;
        }
        if (!hard) {      /* the production path's spring-damper limits */
            R lo = (R)dp->lower[d], hi = (R)dp->upper[d], kl = (R)sp->joint_limit_stiffness, dl = (R)sp->joint_limit_damping;
            R qp = q[d] + h * qd[d];
            R ref = 0; int on = 0;
            if (lo > -1e30f && (q[d] < lo || qp < lo)) { ref = lo; on = 1; }
            else if (hi < 1e30f && (q[d] > hi || qp > hi)) { ref = hi; on = 1; }
            if (on) {
                tau[d] += kl * (ref - q[d]) - (dl + h * kl) * qd[d];
                dext[d] += h * dl + h * h * kl;
            }
        }
    }
    FN(orc_kinematics)(m, root13, q, qd, k);
    if (FN(orc_aba_backward)(m, k, tau, dext) != 0) { free(k); free(con); return -2; }
    FN(orc_aba_forward)(m, k, grav, qdd, a0);
    R v0[6] = {0, 0, 0, 0, 0, 0};
    if (!m->fixed_base) {
        R w[3] = {k->vel[0][0], k->vel[0][1], k->vel[0][2]}, vl[3] = {k->vel[0][3], k->vel[0][4], k->vel[0][5]}, wxv[3];
        FN(cross3)(w, vl, wxv);
        for (int i = 0; i < 3; i++) { v0[i] = w[i] + h * a0[i]; v0[3 + i] = vl[i] + h * (a0[3 + i] + wxv[i]); }
    }
    for (int d = 0; d < nd; d++) qd[d] += h * qdd[d];

    /* ---- every candidate inside the contact offset ---- */
    int nc = 0;
    R mu_g = hf && hfs ? (R)hf->friction : (R)sp->plane_dynamic_friction;
    R mu = (R)0.5 * (mu_g + mu_shape);
    int ground = (hf && hfs) || sp->has_ground;
    R box_c[3], box_h[3];
    FN(orc_root_box)(m, box_c, box_h);
    for (int i = 0; i < m->n_cpts; i++) {
        int l = m->cp_link[i];
        for (int pass = 0; pass < 2; pass++) {
            FN(orc_contact)* cc = &con[nc];
            if (pass == 0) {
                if (!ground || (m->fixed_base && l == 0)) continue;
                R lp[3] = {m->cp_pos[i][0], m->cp_pos[i][1], m->cp_pos[i][2]}, rc[3], gh, n[3];
                FN(matvec3)(k->rot[l], lp, rc);
                for (int a = 0; a < 3; a++) rc[a] += k->pos[l][a];
                FN(orc_ground)(hf, hfs, root13[0] + rc[0], root13[1] + rc[1], &gh, n);
                R gap = (root13[2] + rc[2] - gh) * n[2] - (R)m->cp_radius[i];
                if (gap >= (R)sp->contact_offset) continue;
                cc->link = l; cc->body = m->cp_body[i]; cc->gap = gap; cc->self_c = 0;
                for (int a = 0; a < 3; a++) { cc->n[a] = n[a]; cc->r[a] = rc[a] - (R)m->cp_radius[i] * n[a]; }
            } else {
                if (!(sp->self_collision && l > 0 && l - 1 - m->chain_start[m->cp_chain[i]] >= 1)) continue;
                if (!FN(orc_self_candidate)(m, k, i, (R)sp->contact_offset, box_c, box_h, cc)) continue;
            }
            nc++;
            FN(orc_finish_contact)(m, k, cc);
        }
    }
    /* ---- hard joint-limit rows: sign * qd >= -gap / h ---- */
    int nlim = 0, lim_d[2 * B2G_MAX_DOF];
    R lim_sign[2 * B2G_MAX_DOF], lim_gap[2 * B2G_MAX_DOF], lim_W[2 * B2G_MAX_DOF], lim_lam[2 * B2G_MAX_DOF];
    for (int d = 0; d < nd && hard; d++) {
        R lo = (R)dp->lower[d], hi = (R)dp->upper[d];
        R dv0[6] = {0, 0, 0, 0, 0, 0}, dqd[B2G_MAX_DOF];
        if (!(lo > -1e30f) && !(hi < 1e30f)) continue;
        for (int e = 0; e < nd; e++) dqd[e] = 0;
        FN(orc_apply_impulse)(m, k, 0, 0, d, (R)1, dv0, dqd);
        if (lo > -1e30f) { lim_d[nlim] = d; lim_sign[nlim] = 1; lim_gap[nlim] = q[d] - lo; lim_W[nlim] = dqd[d]; lim_lam[nlim] = 0; nlim++; }
        if (hi < 1e30f) { lim_d[nlim] = d; lim_sign[nlim] = -1; lim_gap[nlim] = hi - q[d]; lim_W[nlim] = dqd[d]; lim_lam[nlim] = 0; nlim++; }
    }

    R maxdep = (R)sp->max_depenetration_velocity;
    R vpos0[6], qdpos[B2G_MAX_DOF];
    int capped = 0;
    info[1] = info[2] = 0;
    for (int phase = 0; phase < 2; phase++) {
        int with_bias = phase == 0;
        int skip = with_bias ? sp->num_position_iterations <= 0 : sp->num_velocity_iterations <= 0;
        int it = 0;
        for (; it < max_iter && !skip && (nc > 0 || nlim > 0); it++) {
            /* convergence is judged on the VELOCITIES (root + joints) between the ends of two sweeps: with redundant contacts (four
             * coplanar corners on one body) the impulse distribution is not unique and may keep drifting along the null space of the
             * Delassus matrix while the motion and the net force per body have long converged */
            R maxd = 0, maxl = 0, vprev[6 + B2G_MAX_DOF];
            for (int i = 0; i < 6; i++) vprev[i] = v0[i];
            for (int d = 0; d < nd; d++) vprev[6 + d] = qd[d];
            for (int s = 0; s < nc; s++) {
                FN(orc_contact)* cc = &con[s];
                R pv[3], F[6];
                FN(orc_contact_velocity)(m, k, cc, v0, qd, pv);
                R vn = pv[0] * cc->n[0] + pv[1] * cc->n[1] + pv[2] * cc->n[2];
                R vt1 = pv[0] * cc->t1[0] + pv[1] * cc->t1[1] + pv[2] * cc->t1[2];
                R vt2 = pv[0] * cc->t2[0] + pv[1] * cc->t2[1] + pv[2] * cc->t2[2];
                R tgt = -cc->gap / h;
                if (tgt > maxdep) tgt = maxdep;
                if (!with_bias && tgt > 0) tgt = 0;
                R dl[3];
                R ln = cc->lam[0] - (vn - tgt) / cc->A[0];
                if (ln < 0) ln = 0;
                dl[0] = ln - cc->lam[0];
                vt1 += cc->A[3] * dl[0]; vt2 += cc->A[6] * dl[0];
                /* EXACT Coulomb solve of this contact's tangential rows for the normal impulse just computed: with b = the tangential
                 * velocity the contact would have with zero tangential impulse, either the sticking impulse -A_tt^-1 b lies inside the
                 * disc of radius mu * lambda_n, or the contact slides and lambda_t = -(A_tt + s I)^-1 b with s > 0 such that
                 * |lambda_t| = mu * lambda_n -- friction opposite to the resulting sliding velocity (v_t = s lambda_t).  s by Newton on
                 * 1/|lambda_t(s)| - 1/(mu lambda_n) (monotone from s = 0).  The production path reaches the same fixed point with a
                 * cheaper update (block solve when sticking, scalar proximal step when sliding). */
                R a = cc->A[4], c2 = cc->A[7], d2 = cc->A[8];
                R b1 = vt1 - (a * cc->lam[1] + c2 * cc->lam[2]), b2 = vt2 - (c2 * cc->lam[1] + d2 * cc->lam[2]);
                R lim_t = mu * ln;
                R det = a * d2 - c2 * c2;
                R l1 = -(d2 * b1 - c2 * b2) / det, l2 = -(a * b2 - c2 * b1) / det;
                if ((R)sqrt((double)(l1 * l1 + l2 * l2)) > lim_t) {
                    if (lim_t <= 0) { l1 = 0; l2 = 0; }
                    else {
                        R sft = 0;
                        for (int nit = 0; nit < 60; nit++) {
                            R aa = a + sft, dd = d2 + sft, dt2 = aa * dd - c2 * c2;
                            R p1 = -(dd * b1 - c2 * b2) / dt2, p2 = -(aa * b2 - c2 * b1) / dt2;
                            R pn = (R)sqrt((double)(p1 * p1 + p2 * p2));
                            l1 = p1; l2 = p2;
                            if (!(pn > 0)) break;
                            R z1 = (dd * p1 - c2 * p2) / dt2, z2 = (aa * p2 - c2 * p1) / dt2;      /* (A_tt + s I)^-1 p */
                            R phi = 1 / pn - 1 / lim_t, dphi = (p1 * z1 + p2 * z2) / (pn * pn * pn);
                            if (phi > -(R)1e-14 / lim_t && phi < (R)1e-14 / lim_t) break;
                            R step = -phi / dphi;
                            sft += step;
                            if (sft < 0) sft = 0;
                            if ((step < 0 ? -step : step) <= (R)1e-15 * (1 + sft)) break;
                        }
                        R pn = (R)sqrt((double)(l1 * l1 + l2 * l2));
                        if (pn > 0) { l1 *= lim_t / pn; l2 *= lim_t / pn; }
                    }
                }
                dl[1] = l1 - cc->lam[1]; dl[2] = l2 - cc->lam[2];
                cc->lam[0] = ln; cc->lam[1] = l1; cc->lam[2] = l2;
                for (int a = 0; a < 3; a++) {
                    R ad = dl[a] < 0 ? -dl[a] : dl[a], al = cc->lam[a] < 0 ? -cc->lam[a] : cc->lam[a];
                    if (ad > maxd) maxd = ad;
                    if (al > maxl) maxl = al;
                }
                R dir[3];
                for (int a = 0; a < 3; a++) dir[a] = cc->n[a] * dl[0] + cc->t1[a] * dl[1] + cc->t2[a] * dl[2];
                FN(orc_contact_wrench)(cc, dir, F);
                FN(orc_apply_contact_impulse)(m, k, cc, F, v0, qd);
            }
            for (int r = 0; r < nlim; r++) {
                int d = lim_d[r];
                R v = lim_sign[r] * qd[d];
                R tgt = -lim_gap[r] / h;
                if (!with_bias && tgt > 0) tgt = 0;
                R ln = lim_lam[r] - (v - tgt) / lim_W[r];
                if (ln < 0) ln = 0;
                R dl = ln - lim_lam[r];
                lim_lam[r] = ln;
                R ad = dl < 0 ? -dl : dl;
                if (ad > maxd) maxd = ad;
                if (ln > maxl) maxl = ln;
                if (dl != 0) FN(orc_apply_impulse)(m, k, 0, 0, d, lim_sign[r] * dl, v0, qd);
            }
            (void)maxl;
            maxd = 0;
            R vmax = 0;
            for (int i = 0; i < 6 + nd; i++) {
                R vn_ = i < 6 ? v0[i] : qd[i - 6];
                R ad = vn_ - vprev[i]; if (ad < 0) ad = -ad;
                R av = vn_ < 0 ? -vn_ : vn_;
                if (ad > maxd) maxd = ad;
                if (av > vmax) vmax = av;
            }
            if (maxd <= tol * (1 + vmax)) { it++; break; }
        }
        if (it >= max_iter && !skip && (nc > 0 || nlim > 0)) capped = 1;
        info[1 + phase] = it;
        if (with_bias) {
            for (int i = 0; i < 6; i++) vpos0[i] = v0[i];
            for (int d = 0; d < nd; d++) qdpos[d] = qd[d];
        }
    }
    if (sp->num_velocity_iterations <= 0 || sp->num_position_iterations <= 0) {
        /* single-phase configurations: positions integrate with the final velocity when there is no velocity phase */
        if (sp->num_velocity_iterations <= 0) { for (int i = 0; i < 6; i++) vpos0[i] = v0[i]; for (int d = 0; d < nd; d++) qdpos[d] = qd[d]; }
    }
    info[0] = nc; info[3] = capped;

    if (!m->fixed_base) {   /* root velocity limits (b2g_sim_params::max_linear_velocity / max_angular_velocity; 0 = none) */
        R* vs[2] = {v0, vpos0};
        for (int k = 0; k < 2; k++) {
            R* v = vs[k];
            R lim_w = (R)sp->max_angular_velocity, lim_v = (R)sp->max_linear_velocity;
            R w2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], l2 = v[3] * v[3] + v[4] * v[4] + v[5] * v[5];
            if (lim_w > 0 && w2 > lim_w * lim_w) { R sc = lim_w / (R)sqrt((double)w2); v[0] *= sc; v[1] *= sc; v[2] *= sc; }
            if (lim_v > 0 && l2 > lim_v * lim_v) { R sc = lim_v / (R)sqrt((double)l2); v[3] *= sc; v[4] *= sc; v[5] *= sc; }
        }
    }
    for (int d = 0; d < nd; d++) {
        R vl = (R)dp->velocity[d];
        if (vl > 0) { if (qd[d] > vl) qd[d] = vl; if (qd[d] < -vl) qd[d] = -vl; if (qdpos[d] > vl) qdpos[d] = vl; if (qdpos[d] < -vl) qdpos[d] = -vl; }
        q[d] += h * qdpos[d];
        dof[2 * d] = q[d]; dof[2 * d + 1] = qd[d];
    }
    if (!m->fixed_base) {
        for (int i = 0; i < 3; i++) root13[i] += h * vpos0[3 + i];
        R w[3] = {vpos0[0], vpos0[1], vpos0[2]};
        R ang = (R)sqrt((double)(w[0] * w[0] + w[1] * w[1] + w[2] * w[2])) * h;
        R dq[4] = {0, 0, 0, 1};
        if (ang > (R)1e-12) {
            R s = (R)sin((double)(ang / 2)) / (ang / h);
            dq[0] = w[0] * s; dq[1] = w[1] * s; dq[2] = w[2] * s; dq[3] = (R)cos((double)(ang / 2));
        }
        R* qo = root13 + 3;
        R x1 = dq[0], y1 = dq[1], z1 = dq[2], w1 = dq[3], x2 = qo[0], y2 = qo[1], z2 = qo[2], w2 = qo[3];
        R nq[4] = {w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                   w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2};
        R nn = 1 / (R)sqrt((double)(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]));
        for (int i = 0; i < 4; i++) qo[i] = nq[i] * nn;
        for (int i = 0; i < 3; i++) { root13[7 + i] = v0[3 + i]; root13[10 + i] = v0[i]; }
    }
    for (int d = 0; d < nd; d++) {
        R f = 0, lim_e = (R)dp->effort[d];
        if (dp->drive_mode[d] == B2G_DOF_MODE_POS) f = (R)dp->stiffness[d] * (target[d] - q[d]) - (R)dp->damping[d] * qd[d];
        else if (dp->drive_mode[d] == B2G_DOF_MODE_VEL) f = (R)dp->damping[d] * (target[d] - qd[d]);
        else if (dp->drive_mode[d] == B2G_DOF_MODE_EFFORT) f = actuation[d];
        if (lim_e > 0) { if (f > lim_e) f = lim_e; if (f < -lim_e) f = -lim_e; }
        dof_force[d] = f;
    }
    for (int b = 0; b < m->n_bodies * 3; b++) contact[b] = 0;
    for (int s = 0; s < nc; s++) {
        FN(orc_contact)* cc = &con[s];
        for (int a = 0; a < 3; a++) {
            R f = (cc->n[a] * cc->lam[0] + cc->t1[a] * cc->lam[1] + cc->t2[a] * cc->lam[2]) / h;
            contact[cc->body * 3 + a] += f;
            if (cc->self_c) contact[0 * 3 + a] -= f;
        }
    }
    free(k); free(con);
    return 0;
}

/* gym.simulate with the converged reference solver.  info: (n_envs, 4) ints of the LAST sub-step of each environment. */
int FN(orc_simulate_ref)(const b2g_model* m, const b2g_sim_params* sp, const b2g_dof_props* dp,
                         const b2g_heightfield* hf, const int16_t* hfs, const float* friction, int n_envs,
                         R* root, R* dof, const R* target, const R* actuation, R* dof_force, R* contact,
                         int flags, int max_iter, double tol, int* info) {
    int nd = m->n_dof, nb = m->n_bodies;
    int nsub = sp->substeps > 0 ? sp->substeps : 1;
    R h = (R)sp->dt / (R)nsub;
    int rc = 0;
    for (int e = 0; e < n_envs; e++) {
        R dummy_root[13] = {0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0};
        R* r = root ? root + (size_t)e * 13 : dummy_root;
        int inf[4] = {0, 0, 0, 0}, worst = 0;
        for (int s = 0; s < nsub; s++) {
            int st = FN(orc_substep_ref)(m, sp, dp, hf, hfs, friction ? (R)friction[e] : (R)1, h, r, dof + (size_t)e * nd * 2,
                                         target + (size_t)e * nd, actuation + (size_t)e * nd, dof_force + (size_t)e * nd,
                                         contact + (size_t)e * nb * 3, flags, max_iter, (R)tol, inf);
            if (st != 0) rc = st;
            worst |= inf[3];
        }
        if (info) { info[e * 4] = inf[0]; info[e * 4 + 1] = inf[1]; info[e * 4 + 2] = inf[2]; info[e * 4 + 3] = worst; }
    }
    return rc;
}

/* forward dynamics probe: joint accelerations + root spatial acceleration (contact-free, no drives) */
int FN(orc_forward_dynamics)(const b2g_model* m, const b2g_sim_params* sp, int n_envs, const R* root, const R* dof,
                             const R* tau, R* qdd, R* root_acc) {
    int nd = m->n_dof;
    FN(orc_kin)* k = (FN(orc_kin)*)malloc(sizeof(FN(orc_kin)));
    if (!k) return -1;
    R grav[3] = {(R)sp->gravity[0], (R)sp->gravity[1], (R)sp->gravity[2]};
    R dext[B2G_MAX_DOF];
    for (int d = 0; d < nd; d++) dext[d] = (R)m->armature[d];
    int rc = 0;
    for (int e = 0; e < n_envs; e++) {
        R q[B2G_MAX_DOF], qd[B2G_MAX_DOF];
        R dummy_root[13] = {0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0};
        const R* r = root ? root + (size_t)e * 13 : dummy_root;
        for (int d = 0; d < nd; d++) { q[d] = dof[((size_t)e * nd + d) * 2]; qd[d] = dof[((size_t)e * nd + d) * 2 + 1]; }
        FN(orc_kinematics)(m, r, q, qd, k);
        if (FN(orc_aba_backward)(m, k, tau + (size_t)e * nd, dext) != 0) rc = -2;
        FN(orc_aba_forward)(m, k, grav, qdd + (size_t)e * nd, root_acc + (size_t)e * 6);
    }
    free(k);
    return rc;
}

/* CRBA mass matrix H ((6+nd)^2 floating: root block first (angular, linear), or nd^2 fixed) and RNEA bias
 * C (same ordering) with gravity, so that H [a0; qdd] + C = [0; tau]. Independent of the ABA code path. */
int FN(orc_crba_rnea)(const b2g_model* m, const b2g_sim_params* sp, const R* root13, const R* dofs, R* H, R* C) {
    int nd = m->n_dof, nb = m->fixed_base ? 0 : 6, n = nd + nb;
    FN(orc_kin)* k = (FN(orc_kin)*)malloc(sizeof(FN(orc_kin)));
    if (!k) return -1;
    R q[B2G_MAX_DOF], qd[B2G_MAX_DOF];
    R dummy_root[13] = {0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0};
    if (!root13) root13 = dummy_root;
    for (int d = 0; d < nd; d++) { q[d] = dofs[2 * d]; qd[d] = dofs[2 * d + 1]; }
    FN(orc_kinematics)(m, root13, q, qd, k);
    /* composite inertias */
    R Ic[B2G_MAX_LINKS][36];
    for (int l = 0; l <= nd; l++) for (int i = 0; i < 36; i++) Ic[l][i] = k->I[l][i];
    for (int c = 0; c < m->n_chains; c++)
        for (int j = m->chain_len[c] - 1; j >= 0; j--) {
            int l = m->chain_start[c] + j + 1, p = (j == 0) ? 0 : l - 1;
            for (int i = 0; i < 36; i++) Ic[p][i] += Ic[l][i];
        }
    for (int i = 0; i < n * n; i++) H[i] = 0;
    if (nb) for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) H[a * n + b] = Ic[0][a * 6 + b];
    for (int c = 0; c < m->n_chains; c++)
        for (int j = 0; j < m->chain_len[c]; j++) {
            int d = m->chain_start[c] + j;
            R F[6];
            FN(mv6)(Ic[d + 1], k->S[d], F);
            H[(nb + d) * n + nb + d] = FN(dot6)(k->S[d], F) + (R)m->armature[d];
            for (int jj = 0; jj < j; jj++) {
                int e = m->chain_start[c] + jj;
                R v = FN(dot6)(k->S[e], F);
                H[(nb + d) * n + nb + e] = v; H[(nb + e) * n + nb + d] = v;
            }
            if (nb) for (int a = 0; a < 6; a++) { H[a * n + nb + d] = F[a]; H[(nb + d) * n + a] = F[a]; }
        }
    /* RNEA with zero accelerations (relative to gravity field) */
    R grav[3] = {(R)sp->gravity[0], (R)sp->gravity[1], (R)sp->gravity[2]};
    R acc[B2G_MAX_LINKS][6], f[B2G_MAX_LINKS][6];
    for (int i = 0; i < 3; i++) { acc[0][i] = 0; acc[0][3 + i] = -grav[i]; }
    for (int c = 0; c < m->n_chains; c++)
        for (int j = 0; j < m->chain_len[c]; j++) {
            int d = m->chain_start[c] + j, l = d + 1, p = (j == 0) ? 0 : l - 1;
            for (int i = 0; i < 6; i++) acc[l][i] = acc[p][i] + k->cb[d][i];
        }
    for (int l = 0; l <= nd; l++) {
        R ia[6], iv[6], t[6];
        FN(mv6)(k->I[l], acc[l], ia);
        FN(mv6)(k->I[l], k->vel[l], iv);
        FN(crf)(k->vel[l], iv, t);
        for (int i = 0; i < 6; i++) f[l][i] = ia[i] + t[i];
    }
    for (int c = 0; c < m->n_chains; c++)
        for (int j = m->chain_len[c] - 1; j >= 0; j--) {
            int d = m->chain_start[c] + j, l = d + 1, p = (j == 0) ? 0 : l - 1;
            C[nb + d] = FN(dot6)(k->S[d], f[l]);
            for (int i = 0; i < 6; i++) f[p][i] += f[l][i];
        }
    if (nb) for (int i = 0; i < 6; i++) C[i] = f[0][i];
    free(k);
    return 0;
}

/* total kinetic energy, potential energy, and spatial momentum about the WORLD origin (6) */
int FN(orc_energy_momentum)(const b2g_model* m, const b2g_sim_params* sp, const R* root13, const R* dofs, R* ke, R* pe, R* mom) {
    int nd = m->n_dof;
    FN(orc_kin)* k = (FN(orc_kin)*)malloc(sizeof(FN(orc_kin)));
    if (!k) return -1;
    R q[B2G_MAX_DOF], qd[B2G_MAX_DOF];
    R dummy_root[13] = {0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0};
    if (!root13) root13 = dummy_root;
    for (int d = 0; d < nd; d++) { q[d] = dofs[2 * d]; qd[d] = dofs[2 * d + 1]; }
    FN(orc_kinematics)(m, root13, q, qd, k);
    R T = 0, V = 0, hO[6] = {0, 0, 0, 0, 0, 0};
    for (int l = 0; l <= nd; l++) {
        R iv[6], cw[3], com[3] = {m->link_com[l][0], m->link_com[l][1], m->link_com[l][2]};
        FN(mv6)(k->I[l], k->vel[l], iv);
        T += (R)0.5 * FN(dot6)(k->vel[l], iv);
        for (int i = 0; i < 6; i++) hO[i] += iv[i];
        FN(matvec3)(k->rot[l], com, cw);
        for (int i = 0; i < 3; i++) V -= (R)m->link_mass[l] * (R)sp->gravity[i] * (root13[i] + k->pos[l][i] + cw[i]);
    }
    /* shift the moment from O (root origin) to the world origin: n_w = n_O + p0 x f */
    R sh[3];
    FN(cross3)(root13, hO + 3, sh);
    for (int i = 0; i < 3; i++) { mom[i] = hO[i] + sh[i]; mom[3 + i] = hO[3 + i]; }
    *ke = T; *pe = V;
    free(k);
    return 0;
}

#undef R
#undef FN
#undef ORC_CAT
#undef ORC_CAT2
