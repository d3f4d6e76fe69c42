/*
 * sph_poiseuille_main.c -- standalone driver of the B200 particle loop for hosts without MATLAB.
 *
 * Restates the thin slice of SPH_Poiseuille.m that sits around the two MEX files, in plain C over
 * the C ABI of libsph_b200 (include/sph_b200.h); every particle-loop operation runs on the GPU.
 *
 *   S2  config.ini grammar and keys          SPH_Poiseuille.m:43-59, 447-512 (config.ini:4-19)
 *       derived parameters                   :62-80
 *   S3  lattice initial condition            :95-123, build_shell_wall_particles.m:14-38
 *   S4  restart (own flat binary format; the reference's is MATLAB HDF5)   :127-163, 434-445, 514-517
 *   S5  first list, density + KGC, EOS       :165-173
 *   S6  main loop, log lines, output points  :246-302
 *   S7  full-length profile and L2           :617-623, SPH_Poiseuille_postprocess.m:42-48
 *
 * Usage: sph_poiseuille_b200 [config.ini] [--dry-run] [--max-steps N] [--dual N_INNER] [--device D] [--slabs N]
 *   --slabs N   the channel sharded into N x-slabs over the visible GPUs (sphb_slabs_*: peer-to-peer mailboxes,
 *               one CUDA graph launch per GPU and step; slabs share GPUs when there are fewer than N)
 * Environment overrides as in the reference (m:19-26): SPH_CONFIG_OVERRIDE, SPH_RESTART_PATH_OVERRIDE;
 * SPH_PROFILE_CSV_OVERRIDE names the CSV that replaces the postprocess MAT file.
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sph_b200.h"

/* ------------------------------------------------------------------------------------------- */
/* INI: "[section]", "key = value  # comment", ';' or '#' start a comment (m:447-499)              */
/* ------------------------------------------------------------------------------------------- */
#define MAX_KEYS 64
typedef struct {
    char section[64], key[64], raw[128];
    double num;
    int is_num;
} ini_entry;
typedef struct {
    ini_entry e[MAX_KEYS];
    int n;
} ini_t;

static char *trim(char *s)
{
    char *e;
    while (*s && isspace((unsigned char)*s)) ++s;
    e = s + strlen(s);
    while (e > s && isspace((unsigned char)e[-1])) --e;
    *e = '\0';
    return s;
}

static int is_varname(const char *s)   /* MATLAB isvarname: letter, then letters/digits/underscores */
{
    if (!isalpha((unsigned char)*s)) return 0;
    for (; *s; ++s)
        if (!isalnum((unsigned char)*s) && *s != '_') return 0;
    return 1;
}

static int ini_parse(const char *path, ini_t *ini)
{
    FILE *f = fopen(path, "r");
    char line[1024], section[64] = "";
    if (!f) {
        fprintf(stderr, "config file not found: %s\n", path);
        return 1;
    }
    ini->n = 0;
    while (fgets(line, sizeof(line), f)) {
        char *s = trim(line), *eq, *key, *val, *c, *end;
        size_t len = strlen(s);
        if (!len || *s == ';' || *s == '#') continue;
        if (*s == '[' && s[len - 1] == ']') {
            s[len - 1] = '\0';
            snprintf(section, sizeof(section), "%s", trim(s + 1));
            if (!is_varname(section)) {
                fprintf(stderr, "illegal INI section name: %s\n", section);
                fclose(f);
                return 1;
            }
            continue;
        }
        eq = strchr(s, '=');
        if (!eq) continue;
        *eq = '\0';
        key = trim(s);
        val = eq + 1;
        for (c = val; *c; ++c)
            if (*c == ';' || *c == '#') {   /* regexprep(val, '[;#].*$', ''), m:483 */
                *c = '\0';
                break;
            }
        val = trim(val);
        if (!section[0] || !is_varname(key) || ini->n >= MAX_KEYS) {
            fprintf(stderr, "bad INI line (key '%s')\n", key);
            fclose(f);
            return 1;
        }
        snprintf(ini->e[ini->n].section, 64, "%s", section);
        snprintf(ini->e[ini->n].key, 64, "%s", key);
        snprintf(ini->e[ini->n].raw, 128, "%s", val);
        ini->e[ini->n].num = strtod(val, &end);   /* str2double: the whole string must be a number */
        ini->e[ini->n].is_num = (end != val && *trim(end) == '\0');
        ini->n++;
    }
    fclose(f);
    return 0;
}

static int ini_num(const ini_t *ini, const char *sec, const char *key, double *out, int required)
{
    int k, found = -1;
    for (k = 0; k < ini->n; ++k)   /* a later duplicate overwrites, like the struct assignment at m:494 */
        if (!strcmp(ini->e[k].section, sec) && !strcmp(ini->e[k].key, key)) found = k;
    if (found < 0) {
        if (required) fprintf(stderr, "missing key: [%s].%s\n", sec, key);
        return required ? 1 : -1;
    }
    if (!ini->e[found].is_num || !isfinite(ini->e[found].num)) {
        fprintf(stderr, "key [%s].%s is not a valid number\n", sec, key);
        return 1;
    }
    *out = ini->e[found].num;
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
typedef struct {
    double DL, DH, dp, rho0, mu, U_bulk, c_f, t_end, output_interval;
    int sort_interval, restart_from_file;
    double transport_coeff, h_over_dp, wall_layers;   /* hard-coded in the script; optional keys here */
    double gravity_g, U_max, h, wall_thickness, p0, inv_sigma0, nu;
    char signature[512];
} cfg_t;

static double matlab_round(double v) { return v < 0 ? -floor(-v + 0.5) : floor(v + 0.5); }

static int load_config(const char *path, cfg_t *c)
{
    ini_t ini;
    double v;
    if (ini_parse(path, &ini)) return 1;
#define REQ(sec, key, dst)                           \
    if (ini_num(&ini, sec, key, &v, 1)) return 1;    \
    dst = v
    REQ("physical", "DL", c->DL);
    REQ("physical", "DH", c->DH);
    REQ("physical", "dp", c->dp);
    REQ("physical", "rho0", c->rho0);
    REQ("physical", "mu", c->mu);
    REQ("physical", "U_bulk", c->U_bulk);
    REQ("physical", "c_f", c->c_f);
    REQ("simulation", "end_time", c->t_end);
    REQ("simulation", "output_interval", c->output_interval);
    REQ("simulation", "sort_interval", v);
    c->sort_interval = (int)matlab_round(v);
    REQ("simulation", "restart_from_file", v);
    c->restart_from_file = (int)matlab_round(v);
#undef REQ
    c->transport_coeff = 0.30;   /* m:77 */
    c->h_over_dp = 1.3;          /* m:73 */
    c->wall_layers = 4.0;        /* m:74 */
    if (ini_num(&ini, "numerics", "transport_coeff", &v, 0) == 0) c->transport_coeff = v;
    if (ini_num(&ini, "numerics", "h_over_dp", &v, 0) == 0) c->h_over_dp = v;
    if (ini_num(&ini, "numerics", "wall_layers", &v, 0) == 0) c->wall_layers = v;
    if (!(c->dp > 0) || !(c->DL > 0) || !(c->DH > 0) || !(c->rho0 > 0)) {
        fprintf(stderr, "DL, DH, dp and rho0 must be positive\n");
        return 1;
    }
    c->DL = matlab_round(c->DL / c->dp) * c->dp;                                   /* m:64 */
    c->DH = matlab_round(c->DH / c->dp) * c->dp;                                   /* m:65 */
    c->gravity_g = 12.0 * c->mu * c->U_bulk / (c->rho0 * c->DH * c->DH);           /* m:71 */
    c->U_max = 1.5 * c->U_bulk;                                                    /* m:72 */
    c->h = c->h_over_dp * c->dp;                                                   /* m:73 */
    c->wall_thickness = fmax(c->wall_layers * c->dp, ceil(2.0 * c->h / c->dp) * c->dp);   /* m:74-75 */
    c->p0 = c->rho0 * c->c_f * c->c_f;                                             /* m:78 */
    c->inv_sigma0 = c->dp * c->dp;                                                 /* m:79 */
    c->nu = c->mu / c->rho0;                                                       /* m:80 */
    snprintf(c->signature, sizeof(c->signature),                                   /* m:514-517 */
             "DL=%.12g|DH=%.12g|dp=%.12g|rho0=%.12g|mu=%.12g|Ub=%.12g|cf=%.12g|t=%.12g|oi=%.12g|si=%d|"
             "wall=thick-wall-noslip-dual-dt",
             c->DL, c->DH, c->dp, c->rho0, c->mu, c->U_bulk, c->c_f, c->t_end, c->output_interval, c->sort_interval);
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
typedef struct {
    int n_fluid, n_total;
    double *pos, *vel, *rho, *drho_dt, *mass, *wall_vel;   /* column-major [n_total x k] */
} particles_t;

static int build_lattice(const cfg_t *c, particles_t *p)
{
    const int nx = (int)matlab_round(c->DL / c->dp), ny = (int)matlab_round(c->DH / c->dp);
    const int nl = (int)matlab_round(c->wall_thickness / c->dp);
    const int nf = nx * ny, nt = nf + 2 * nx * nl;
    const double y0[3] = {c->dp / 2, -c->wall_thickness + c->dp / 2, c->DH + c->dp / 2};
    const int cnt[3] = {ny, nl, nl};
    int b, ix, iy, k = 0;
    p->n_fluid = nf;
    p->n_total = nt;
    p->pos = (double *)calloc((size_t)2 * nt, sizeof(double));
    p->vel = (double *)calloc((size_t)2 * nt, sizeof(double));
    p->wall_vel = (double *)calloc((size_t)2 * nt, sizeof(double));
    p->rho = (double *)calloc((size_t)nt, sizeof(double));
    p->drho_dt = (double *)calloc((size_t)nt, sizeof(double));
    p->mass = (double *)calloc((size_t)nt, sizeof(double));
    if (!p->pos || !p->vel || !p->wall_vel || !p->rho || !p->drho_dt || !p->mass) return 1;
    /* fluid (meshgrid(x,y)(:): y fastest, m:97-98), bottom wall block, top wall block (y fastest) */
    for (b = 0; b < 3; ++b)
        for (ix = 0; ix < nx; ++ix)
            for (iy = 0; iy < cnt[b]; ++iy, ++k) {
                p->pos[k] = c->dp / 2 + c->dp * ix;
                p->pos[nt + k] = y0[b] + c->dp * iy;
                p->rho[k] = c->rho0;
                p->mass[k] = c->rho0 * c->dp * c->dp;   /* m:117-119 */
            }
    return 0;
}

/* the particle loop behind the driver: one resident context, or the sharded channel (--slabs N) */
typedef struct {
    sphb_ctx *ctx;
    sphb_slabs *grp;
} engine_t;

static int eng_status(engine_t *e, sphb_status *st) { return e->grp ? sphb_slabs_status(e->grp, st) : sphb_ctx_status(e->ctx, st); }
static int eng_profile(engine_t *e, double x0, double half_width, int n_bins, double *u_mean)
{
    return e->grp ? sphb_slabs_profile(e->grp, x0, half_width, n_bins, u_mean, NULL)
                  : sphb_ctx_profile(e->ctx, x0, half_width, n_bins, u_mean, NULL);
}
static int eng_run_until(engine_t *e, double target, double t_end, long long chunk, long long *done)
{
    return e->grp ? sphb_slabs_run_until(e->grp, target, t_end, chunk, done) : sphb_ctx_run_until(e->ctx, target, t_end, chunk, done);
}

/* full-length binned u(y) vs analytic: make_postprocess_data m:617-623 + postprocess:42 */
static double l2_error(const cfg_t *c, engine_t *eng, int n_bins, double *u_mean, double *u_exact, double *y_mid)
{
    double num = 0.0, den = 0.0;
    int k;
    if (eng_profile(eng, -1.0, 0.0, n_bins, u_mean)) return NAN;
    for (k = 0; k < n_bins; ++k) {
        const double w = c->DH / n_bins;   /* linspace(0, DH, n_bins + 1), m:581 */
        const double lo = k * w, hi = (k == n_bins - 1) ? c->DH : (k + 1) * w;
        y_mid[k] = 0.5 * (lo + hi);
        u_exact[k] = c->gravity_g / (2.0 * c->nu) * y_mid[k] * (c->DH - y_mid[k]);
        if (isnan(u_mean[k])) continue;
        num += (u_mean[k] - u_exact[k]) * (u_mean[k] - u_exact[k]);
        den += u_exact[k] * u_exact[k];
    }
    return sqrt(num / fmax(den, 2.220446049250313e-16));
}

/* restart: magic, signature, n_total, t, step, then pos vel rho p drho_dt force force_prior (m:434-445) */
/* format 2: an endianness / version word and n_fluid behind the signature (a file of another byte order, version or
 * fluid / wall split is refused instead of misread).  Not the reference's MAT v7.3 restart: the two codes cannot
 * resume each other's runs. */
static const char RESTART_MAGIC[8] = {'S', 'P', 'H', 'B', 'R', 'S', 'T', '2'};
static const unsigned RESTART_VERSION = 0x01020304u;   /* reads back as 0x04030201 on the other byte order */

static int save_restart(const char *path, const cfg_t *c, engine_t *eng, int nf, int nt)
{
    sphb_status st;
    size_t n = (size_t)nt;
    double *buf = (double *)calloc(11 * n, sizeof(double));
    FILE *f;
    int rc = 1;
    if (!buf) return 1;
    if (eng_status(eng, &st)) goto done;
    /* (the sharded channel carries pos, vel, rho, drho_dt, p; the force blocks are recomputed outputs and stay zero) */
    if (eng->grp ? sphb_slabs_download(eng->grp, buf, buf + 2 * n, buf + 4 * n, buf + 5 * n, buf + 6 * n)
                 : sphb_ctx_download(eng->ctx, buf, buf + 2 * n, buf + 4 * n, buf + 5 * n, buf + 6 * n, buf + 7 * n, buf + 9 * n,
                                     NULL, NULL))
        goto done;
    f = fopen(path, "wb");
    if (!f) goto done;
    {
        unsigned sig_len = (unsigned)strlen(c->signature);
        long long step = st.step;
        rc = !(fwrite(RESTART_MAGIC, 1, 8, f) == 8 && fwrite(&sig_len, sizeof(sig_len), 1, f) == 1 &&
               fwrite(c->signature, 1, sig_len, f) == sig_len && fwrite(&RESTART_VERSION, sizeof(RESTART_VERSION), 1, f) == 1 &&
               fwrite(&nf, sizeof(nf), 1, f) == 1 && fwrite(&nt, sizeof(nt), 1, f) == 1 &&
               fwrite(&st.t, sizeof(double), 1, f) == 1 && fwrite(&step, sizeof(step), 1, f) == 1 &&
               fwrite(buf, sizeof(double), 11 * n, f) == 11 * n);
    }
    fclose(f);
done:
    free(buf);
    return rc;
}

/* returns 0 when a matching restart was loaded (signature and sizes agree, m:132-163) */
static int load_restart(const char *path, const cfg_t *c, particles_t *p, int nt, double *t, long long *step)
{
    FILE *f = fopen(path, "rb");
    char magic[8], sig[512];
    unsigned sig_len = 0;
    int nt_file = 0, nf_file = 0, rc = 1;
    unsigned version = 0;
    size_t n = (size_t)nt;
    double *buf = NULL;
    if (!f) return 1;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, RESTART_MAGIC, 8) || fread(&sig_len, sizeof(sig_len), 1, f) != 1 ||
        sig_len >= sizeof(sig) || fread(sig, 1, sig_len, f) != sig_len)
        goto done;
    sig[sig_len] = '\0';
    if (strcmp(sig, c->signature)) {
        printf("restart signature differs from the current configuration: starting from t = 0\n");
        goto done;
    }
    if (fread(&version, sizeof(version), 1, f) != 1 || version != RESTART_VERSION || fread(&nf_file, sizeof(nf_file), 1, f) != 1 ||
        nf_file != p->n_fluid || fread(&nt_file, sizeof(nt_file), 1, f) != 1 || nt_file != nt || fread(t, sizeof(double), 1, f) != 1 ||
        fread(step, sizeof(*step), 1, f) != 1)
        goto done;
    buf = (double *)malloc(11 * n * sizeof(double));
    if (!buf || fread(buf, sizeof(double), 11 * n, f) != 11 * n) goto done;
    /* carried state is pos, vel, rho, drho_dt; p / force are recomputed outputs.  The rows replace
     * the lattice before the engine is created (the sharded channel is cut from them). */
    memcpy(p->pos, buf, 2 * n * sizeof(double));
    memcpy(p->vel, buf + 2 * n, 2 * n * sizeof(double));
    memcpy(p->rho, buf + 4 * n, n * sizeof(double));
    memcpy(p->drho_dt, buf + 6 * n, n * sizeof(double));
    rc = 0;
done:
    free(buf);
    fclose(f);
    return rc;
}

static const char *env_or(const char *name, const char *fallback)
{
    const char *v = getenv(name);
    return (v && *v) ? v : fallback;
}

#define CHECK(call)                                                                             \
    do {                                                                                        \
        if (call) {                                                                             \
            fprintf(stderr, "%s: %s\n", sphb_last_error_id(), sphb_last_error_msg());           \
            return 2;                                                                           \
        }                                                                                       \
    } while (0)

int main(int argc, char **argv)
{
    const char *config_path = env_or("SPH_CONFIG_OVERRIDE", "config.ini");
    const char *restart_path = env_or("SPH_RESTART_PATH_OVERRIDE", "restart.sphb");
    const char *csv_path = env_or("SPH_PROFILE_CSV_OVERRIDE", "SPH_Poiseuille_profile.csv");
    int dry_run = 0, dual_inner = 0, device = -1, n_slabs = 0, a, k, n_bins;
    long long max_steps = 0, steps_total = 0;
    cfg_t c;
    particles_t p;
    sphb_params prm;
    sphb_ctx *ctx = NULL;
    engine_t eng = {NULL, NULL};
    sphb_status st;
    double t_restart = 0;
    long long step_restart = 0;
    int restarted = 0;
    double tau_target, *u_mean, *u_exact, *y_mid, *mid_profiles = NULL, l2;
    int n_profiles = 0, cap_profiles = 0;
    double *profile_times = NULL;

    for (a = 1; a < argc; ++a) {
        if (!strcmp(argv[a], "--dry-run")) dry_run = 1;
        else if (!strcmp(argv[a], "--max-steps") && a + 1 < argc) max_steps = atoll(argv[++a]);
        else if (!strcmp(argv[a], "--dual") && a + 1 < argc) dual_inner = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--device") && a + 1 < argc) device = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--slabs") && a + 1 < argc) n_slabs = atoi(argv[++a]);
        else if (argv[a][0] != '-') config_path = argv[a];
        else {
            fprintf(stderr, "usage: %s [config.ini] [--dry-run] [--max-steps N] [--dual N_INNER] [--device D] [--slabs N]\n",
                    argv[0]);
            return 1;
        }
    }
    if (load_config(config_path, &c)) return 1;
    if (build_lattice(&c, &p)) return 1;
    tau_target = c.gravity_g * c.rho0 * c.DH / 2;                 /* m:231 */
    n_bins = (int)fmax(20.0, matlab_round(c.DH / c.dp));          /* m:234 */
    printf("config: %s\n", config_path);
    printf("DL=%.12g DH=%.12g dp=%.12g h=%.12g wall=%.12g g=%.12g U_max=%.12g p0=%.12g nu=%.12g\n", c.DL, c.DH, c.dp,
           c.h, c.wall_thickness, c.gravity_g, c.U_max, c.p0, c.nu);
    printf("particles: n_fluid=%d n_wall=%d n_total=%d, n_bins=%d, tau_target=%.6f\n", p.n_fluid, p.n_total - p.n_fluid,
           p.n_total, n_bins, tau_target);
    printf("signature: %s\n", c.signature);
    if (dry_run) return 0;

    memset(&prm, 0, sizeof(prm));
    prm.n_fluid = p.n_fluid;
    prm.n_total = p.n_total;
    prm.DL = c.DL;
    prm.DH = c.DH;
    prm.dp = c.dp;
    prm.h = c.h;
    prm.rho0 = c.rho0;
    prm.p0 = c.p0;
    prm.c_f = c.c_f;
    prm.mu = c.mu;
    prm.gravity_g = c.gravity_g;
    prm.inv_sigma0 = c.inv_sigma0;
    prm.transport_coeff = c.transport_coeff;
    prm.sort_interval = c.sort_interval;
    prm.device = device;
    if (c.restart_from_file) {                                    /* S4 */
        if (load_restart(restart_path, &c, &p, p.n_total, &t_restart, &step_restart) == 0) {
            printf("restart loaded: t=%.6f, step=%lld\n", t_restart, step_restart);
            restarted = 1;
        } else {
            printf("no usable restart at %s: starting from t = 0\n", restart_path);
        }
    }
    if (n_slabs > 1) {
        if (dual_inner > 1) {
            fprintf(stderr, "--dual is not available with --slabs (the sharded channel runs the reference schedule)\n");
            return 1;
        }
        /* the sharded channel: cut, first exchange, first list and shear monitor inside (S5 without the density
         * pass: every step starts with one, m:254) */
        CHECK(sphb_slabs_create(&prm, n_slabs, device >= 0 ? &device : NULL, device >= 0 ? 1 : 0, p.pos, p.vel, p.rho,
                                p.drho_dt, p.mass, p.wall_vel, &eng.grp));
        if (restarted) CHECK(sphb_slabs_set_time(eng.grp, t_restart, step_restart));
        printf("sharded: %d x-slabs", n_slabs);
        for (k = 0; k < n_slabs; ++k) {
            int lo = 0, hi = 0, dev = 0;
            (void)sphb_slabs_member(eng.grp, k, &lo, &hi, &dev);
            printf("%s columns [%d, %d) on GPU %d", k ? "," : ":", lo, hi, dev);
        }
        printf("\n");
    } else {
        CHECK(sphb_ctx_create(&prm, &ctx));
        eng.ctx = ctx;
        CHECK(sphb_ctx_upload(ctx, p.pos, p.vel, p.rho, p.drho_dt, p.mass, p.wall_vel));
        if (restarted) CHECK(sphb_ctx_set_time(ctx, t_restart, step_restart));
        CHECK(sphb_ctx_neighbor_build(ctx));                      /* S5, m:167-173 */
        CHECK(sphb_ctx_density_correction(ctx));
        CHECK(sphb_ctx_eos(ctx));
    }

    u_mean = (double *)malloc(3 * (size_t)n_bins * sizeof(double));
    u_exact = u_mean + n_bins;
    y_mid = u_exact + n_bins;
    CHECK(eng_status(&eng, &st));
    for (;;) {                                                    /* S6 */
        double target;
        /* mid-channel profile at every output point, the initial state included (m:241-244, 297-301) */
        if (n_profiles == cap_profiles) {
            cap_profiles = cap_profiles ? 2 * cap_profiles : 32;
            mid_profiles = (double *)realloc(mid_profiles, (size_t)cap_profiles * n_bins * sizeof(double));
            profile_times = (double *)realloc(profile_times, (size_t)cap_profiles * sizeof(double));
        }
        CHECK(eng_profile(&eng, 0.5 * c.DL, fmax(c.dp, c.h), n_bins, mid_profiles + (size_t)n_profiles * n_bins));
        profile_times[n_profiles++] = st.t;
        if (!(st.t < c.t_end - 1e-12) || (max_steps > 0 && steps_total >= max_steps)) break;
        target = fmin(st.t + c.output_interval, c.t_end);         /* m:251 */
        while (st.t < target - 1e-12 && !(max_steps > 0 && steps_total >= max_steps)) {
            /* 20 steps at a time so the log cadence of m:285-291 is kept */
            long long chunk = 20 - (st.step % 20), done = 0;
            if (max_steps > 0 && chunk > max_steps - steps_total) chunk = max_steps - steps_total;
            if (dual_inner > 1 && st.t + 1.05 * dual_inner * st.dt < target) {
                CHECK(sphb_ctx_run_dual(ctx, 1, dual_inner));
                done = dual_inner;
            } else {
                CHECK(eng_run_until(&eng, target, c.t_end, chunk, &done));
            }
            steps_total += done;
            CHECK(eng_status(&eng, &st));
            if (st.error_flags) {
                fprintf(stderr, "device error flags %d at step %lld\n", st.error_flags, st.step);
                return 3;
            }
            if (st.step % 20 == 0 || dual_inner > 1) {
                printf("step=%lld, t=%.6f/%.6f, dt=%.4e, pairs=%lld, vmax=%.4f\n", st.step, st.t, c.t_end, st.dt,
                       st.n_half_pairs, st.vmax);
                printf("  [thick-wall-noslip] tau_bot=%.4f, tau_top=%.4f, tau_target=%.4f\n", st.tau_bottom, st.tau_top,
                       tau_target);
            }
        }
        printf("output point: t=%.6f, step=%lld\n", st.t, st.step);   /* m:294 */
        if (save_restart(restart_path, &c, &eng, p.n_fluid, p.n_total)) fprintf(stderr, "warning: could not write %s\n", restart_path);
    }

    l2 = l2_error(&c, &eng, n_bins, u_mean, u_exact, y_mid);      /* S7 */
    printf("L2 relative error = %.4f%%\n", 100.0 * l2);
    printf(l2 < 0.05 ? "validation passed: L2_error < 5%%\n" : "validation not met: L2_error >= 5%%\n");
    {
        FILE *f = fopen(csv_path, "w");
        if (f) {
            fprintf(f, "y_mid,u_mean,u_exact");
            for (k = 0; k < n_profiles; ++k) fprintf(f, ",u_mid_t=%.6f", profile_times[k]);
            fprintf(f, "\n");
            for (k = 0; k < n_bins; ++k) {
                int q;
                fprintf(f, "%.17g,%.17g,%.17g", y_mid[k], u_mean[k], u_exact[k]);
                for (q = 0; q < n_profiles; ++q) fprintf(f, ",%.17g", mid_profiles[(size_t)q * n_bins + k]);
                fprintf(f, "\n");
            }
            fclose(f);
            printf("profiles written: %s\n", csv_path);
        }
    }
    if (ctx) {
        sphb_timing tm;
        if (sphb_ctx_timing(ctx, &tm) == 0) printf("kernel launches: %lld\n", tm.launches);
    }
    sphb_ctx_destroy(ctx);
    sphb_slabs_destroy(eng.grp);
    free(u_mean);
    free(mid_profiles);
    free(profile_times);
    return 0;
}
