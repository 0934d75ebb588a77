/* broadcast_b200 -- C ABI of the B200 (sm_100a) implementation of BROADCAST's finite-volume hot path.
 *
 * Drop-in boundary: the reference (onera/Broadcast) exposes its Fortran kernels to Python through
 * f2py modules (srcfv.f_sch, srcfv.f_lin, srcfv.f_bnd, srcfv.f_geom, srcfv.f_norm, misc.f_misc ...).
 * Every `bc_*` function below replaces ONE of those f2py entry points: same name (prefixed), same
 * argument order as the Fortran dummy list, same array conventions:
 *
 *   - all arrays are caller-owned HOST memory, float64 / int32, column-major (Fortran order),
 *     padded shapes  cell arrays (im+2gh, jm+2gh[,5]),  node/face arrays (im+2gh+1, jm+2gh+1[,2]),
 *     Fortran lower bound 1-gh;  `inout` arrays are modified in place;
 *   - `interf`, `prr`, `prd` are int32[4] = {imin, jmin, imax, jmax} (1-based, inclusive: the Fortran
 *     integer(2,2) read as p(1,1), p(1,2), p(2,1), p(2,2), i.e. the C-order image of the numpy array
 *     [[imin,jmin],[imax,jmax]] the drivers build);  `loc` is one of "Ilo","Ihi","Jlo","Jhi";
 *   - return value: 0 on success, otherwise a negative BC_ERR_* code or a positive cudaError_t
 *     (the Fortran has no error path; the Python shim raises on non-zero).
 *
 * The `bcd_*` functions are the same operations on DEVICE pointers (dense layout identical to the
 * host layout), asynchronous on `stream` (a cudaStream_t passed as void*): this is the resident mode
 * used for whole Newton/Jacobian steps without host round trips.
 *
 * There is no CPU fallback: every entry point returns an error if no CUDA device is usable.
 */
#ifndef BROADCAST_B200_H
#define BROADCAST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BC_OK 0
#define BC_ERR_ARG (-1)     /* invalid argument (bad loc string, non-positive size, unsupported gh ...) */
#define BC_ERR_NODEV (-2)   /* no CUDA device */
#define BC_ERR_ALLOC (-3)   /* device allocation failed */
#define BC_ERR_UNSUPPORTED (-4)

/* library / device info */
int bc_version(void);
int bc_device_count(void);
const char* bc_last_error(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
long long bc_launch_count(void);

/* ---- residual: srcfv/rhs/flux_num_dnc5.F90:7-226 (f_sch.flux_num_dnc5_2d),
 *      srcfv/rhs/flux_num_dnc5_nowall.F90:120-157 (f_sch.flux_num_dnc5_nowall_2d).
 *      x0,y0,xc,yc are accepted and ignored (unused by the reference body). residu ghosts untouched. */
int bc_flux_num_dnc5_2d(double* residu, const double* w, const double* x0, const double* y0, const double* nx,
                        const double* ny, const double* xc, const double* yc, const double* vol, const double* volf,
                        int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                        double muref, double tref, double s_suth, double k2, double k4, int im, int jm);
int bc_flux_num_dnc5_nowall_2d(double* residu, const double* w, const double* x0, const double* y0, const double* nx,
                               const double* ny, const double* xc, const double* yc, const double* vol,
                               const double* volf, int gh, double cp, double cv, double prandtl, double gam,
                               double rgaz, double cs, double muref, double tref, double s_suth, double k2, double k4,
                               int im, int jm);

/* ---- tangent: srcfv/tangent/flux_num_dnc5_d.f90:15-3870 (f_lin.flux_num_dnc5_2d_d).
 *      residud is zeroed everywhere then filled on interior cells; residu is NOT written (:3853-3868). */
int bc_flux_num_dnc5_2d_d(double* residu, double* residud, const double* w, const double* wd, const double* x0,
                          const double* y0, const double* nx, const double* ny, const double* xc, const double* yc,
                          const double* vol, const double* volf, int gh, double cp, double cv, double prandtl,
                          double gam, double rgaz, double cs, double muref, double tref, double s_suth, double k2,
                          double k4, int im, int jm);
int bc_flux_num_dnc5_nowall_2d_d(double* residu, double* residud, const double* w, const double* wd, const double* x0,
                                 const double* y0, const double* nx, const double* ny, const double* xc,
                                 const double* yc, const double* vol, const double* volf, int gh, double cp, double cv,
                                 double prandtl, double gam, double rgaz, double cs, double muref, double tref,
                                 double s_suth, double k2, double k4, int im, int jm);

/* ---- the other members of the scheme family: orders 3 / 7 / 9, gh = (order + 1) / 2 = 2 / 4 / 5
 *      (BROADCAST_npz.py:501-502).  Same argument lists and conventions as the order-5 routines above; a routine called with a
 *      ghost depth other than its own returns BC_ERR_UNSUPPORTED.  The resident entry points (bcd_residual, bcd_tangent,
 *      bcd_jacobian_coo, the boundary fills, seeds, scatters, norms) select the member from gh. */
/* order 3: srcfv/rhs/flux_num_dnc3.F90:7-219 (f_sch.flux_num_dnc3_2d), srcfv/rhs/flux_num_dnc3_nowall.F90 (f_sch.flux_num_dnc3_nowall_2d),
 *      srcfv/tangent/flux_num_dnc3_d.f90 and flux_num_dnc3_nowall_d.f90 (f_lin): euler_o4 / predictor_5p / gradop_3p, 2nd-order viscous gradients on every face, wall rows 2 and 1 */
int bc_flux_num_dnc3_2d(double* residu, const double* w, const double* x0, const double* y0, const double* nx,
                        const double* ny, const double* xc, const double* yc, const double* vol, const double* volf,
                        int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                        double muref, double tref, double s_suth, double k2, double k4, int im, int jm);
int bc_flux_num_dnc3_nowall_2d(double* residu, const double* w, const double* x0, const double* y0, const double* nx,
                               const double* ny, const double* xc, const double* yc, const double* vol,
                               const double* volf, int gh, double cp, double cv, double prandtl, double gam,
                               double rgaz, double cs, double muref, double tref, double s_suth, double k2, double k4,
                               int im, int jm);
int bc_flux_num_dnc3_2d_d(double* residu, double* residud, const double* w, const double* wd, const double* x0,
                          const double* y0, const double* nx, const double* ny, const double* xc, const double* yc,
                          const double* vol, const double* volf, int gh, double cp, double cv, double prandtl,
                          double gam, double rgaz, double cs, double muref, double tref, double s_suth, double k2,
                          double k4, int im, int jm);
int bc_flux_num_dnc3_nowall_2d_d(double* residu, double* residud, const double* w, const double* wd, const double* x0,
                                 const double* y0, const double* nx, const double* ny, const double* xc,
                                 const double* yc, const double* vol, const double* volf, int gh, double cp, double cv,
                                 double prandtl, double gam, double rgaz, double cs, double muref, double tref,
                                 double s_suth, double k2, double k4, int im, int jm);
/* order 7: srcfv/rhs/flux_num_dnc7.F90:7-238 (f_sch.flux_num_dnc7_2d), srcfv/rhs/flux_num_dnc7_nowall.F90 (f_sch.flux_num_dnc7_nowall_2d),
 *      srcfv/tangent/flux_num_dnc7_d.f90 and flux_num_dnc7_nowall_d.f90 (f_lin): euler_o8 / predictor_9p / gradop_7p, off-centred wall rows 4, 3, 2 (coefnearbnd_9p.F) */
int bc_flux_num_dnc7_2d(double* residu, const double* w, const double* x0, const double* y0, const double* nx,
                        const double* ny, const double* xc, const double* yc, const double* vol, const double* volf,
                        int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                        double muref, double tref, double s_suth, double k2, double k4, int im, int jm);
int bc_flux_num_dnc7_nowall_2d(double* residu, const double* w, const double* x0, const double* y0, const double* nx,
                               const double* ny, const double* xc, const double* yc, const double* vol,
                               const double* volf, int gh, double cp, double cv, double prandtl, double gam,
                               double rgaz, double cs, double muref, double tref, double s_suth, double k2, double k4,
                               int im, int jm);
int bc_flux_num_dnc7_2d_d(double* residu, double* residud, const double* w, const double* wd, const double* x0,
                          const double* y0, const double* nx, const double* ny, const double* xc, const double* yc,
                          const double* vol, const double* volf, int gh, double cp, double cv, double prandtl,
                          double gam, double rgaz, double cs, double muref, double tref, double s_suth, double k2,
                          double k4, int im, int jm);
int bc_flux_num_dnc7_nowall_2d_d(double* residu, double* residud, const double* w, const double* wd, const double* x0,
                                 const double* y0, const double* nx, const double* ny, const double* xc,
                                 const double* yc, const double* vol, const double* volf, int gh, double cp, double cv,
                                 double prandtl, double gam, double rgaz, double cs, double muref, double tref,
                                 double s_suth, double k2, double k4, int im, int jm);
/* order 9: srcfv/rhs/flux_num_dnc9.F90:7-251 (f_sch.flux_num_dnc9_2d), srcfv/rhs/flux_num_dnc9_nowall.F90 (f_sch.flux_num_dnc9_nowall_2d),
 *      srcfv/tangent/flux_num_dnc9_d.f90 and flux_num_dnc9_nowall_d.f90 (f_lin): euler_o10 / predictor_11p / gradop_9p, off-centred wall rows 5 .. 2 (coefnearbnd_11p.F) */
int bc_flux_num_dnc9_2d(double* residu, const double* w, const double* x0, const double* y0, const double* nx,
                        const double* ny, const double* xc, const double* yc, const double* vol, const double* volf,
                        int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                        double muref, double tref, double s_suth, double k2, double k4, int im, int jm);
int bc_flux_num_dnc9_nowall_2d(double* residu, const double* w, const double* x0, const double* y0, const double* nx,
                               const double* ny, const double* xc, const double* yc, const double* vol,
                               const double* volf, int gh, double cp, double cv, double prandtl, double gam,
                               double rgaz, double cs, double muref, double tref, double s_suth, double k2, double k4,
                               int im, int jm);
int bc_flux_num_dnc9_2d_d(double* residu, double* residud, const double* w, const double* wd, const double* x0,
                          const double* y0, const double* nx, const double* ny, const double* xc, const double* yc,
                          const double* vol, const double* volf, int gh, double cp, double cv, double prandtl,
                          double gam, double rgaz, double cs, double muref, double tref, double s_suth, double k2,
                          double k4, int im, int jm);
int bc_flux_num_dnc9_nowall_2d_d(double* residu, double* residud, const double* w, const double* wd, const double* x0,
                                 const double* y0, const double* nx, const double* ny, const double* xc,
                                 const double* yc, const double* vol, const double* volf, int gh, double cp, double cv,
                                 double prandtl, double gam, double rgaz, double cs, double muref, double tref,
                                 double s_suth, double k2, double k4, int im, int jm);

/* ---- isothermal-wall variant of the scheme (`_iso`): srcfv/rhs/flux_num_dnc5_iso.F90:7-227 =
 *      flux_num_dnc5_2d with rhs/fluxwall_iso.F (heat flux lambda (T - twall) n / vol_face through the wall face, lambda as the
 *      i-face viscous fragment left it: flux_visqueux_o2_i.F:61) and its tangent srcfv/tangent/flux_num_dnc5_iso_d.f90:15-3877
 *      (twall passive).  Same argument lists as the reference: twall follows w (and wd). */
int bc_flux_num_dnc5_iso_2d(double* residu, const double* w, double twall, const double* x0, const double* y0, const double* nx,
                            const double* ny, const double* xc, const double* yc, const double* vol, const double* volf,
                            int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                            double muref, double tref, double s_suth, double k2, double k4, int im, int jm);
int bc_flux_num_dnc5_iso_2d_d(double* residu, double* residud, const double* w, const double* wd, double twall,
                              const double* x0, const double* y0, const double* nx, const double* ny, const double* xc,
                              const double* yc, const double* vol, const double* volf, int gh, double cp, double cv,
                              double prandtl, double gam, double rgaz, double cs, double muref, double tref, double s_suth,
                              double k2, double k4, int im, int jm);

/* ---- boundary fills (f_bnd.*) and their tangents (f_lin.*_d): srcfv/borders/ (.F90), srcfv/tangent/bc_*_d.f90 */
int bc_bc_wall_viscous_adia_2d(double* w, const char* loc, double gam, const int32_t* interf, int gh, int im, int jm);
int bc_bc_wall_viscous_adia_2d_d(double* w, double* wd, const char* loc, double gam, const int32_t* interf, int gh,
                                 int im, int jm);
int bc_bc_no_reflexion_2d(double* w, const double* wbd, const char* loc, const int32_t* interf, const double* nx,
                          const double* ny, double gam, int gh, int im, int jm, int lm);
int bc_bc_no_reflexion_2d_d(double* w, double* wd, const double* wbd, const char* loc, const int32_t* interf,
                            const double* nx, const double* ny, double gam, int gh, int im, int jm, int lm);
int bc_bc_supandsubinlet_2d(double* w, const char* loc, const int32_t* interf, const double* field, const double* nx,
                            const double* ny, double gam, int im, int jm, int lm, int gh);
int bc_bc_supandsubinlet_2d_d(double* w, double* wd, const char* loc, const int32_t* interf, const double* field,
                              const double* nx, const double* ny, double gam, int im, int jm, int lm, int gh);
int bc_bc_extrapolate_o2_2d(double* w, const char* loc, const int32_t* interf, int im, int jm, int gh, int em);
int bc_bc_extrapolate_o2_2d_d(double* w, double* wd, const char* loc, const int32_t* interf, int im, int jm, int gh,
                              int em);
/* isothermal wall and symmetry plane (the sensitivity driver's walls, half-domain cards card_cyl2d.py:105):
 * srcfv/borders/bc_wall_viscous_iso.F90:1-108 + srcfv/tangent/bc_wall_viscous_iso_d.f90; srcfv/borders/bc_symmetry.F90:1-79 +
 * srcfv/tangent/bc_symmetry_d.f90 (call site handleBC.py:228-231: fsym(w, loc, interf, nx, ny, gh, im, jm)) */
int bc_bc_wall_viscous_iso_2d(double* w, double twall, const char* loc, double gam, double rgaz, const int32_t* interf, int gh,
                              int im, int jm);
int bc_bc_wall_viscous_iso_2d_d(double* w, double* wd, double twall, const char* loc, double gam, double rgaz,
                                const int32_t* interf, int gh, int im, int jm);
int bc_bc_symmetry_2d(double* w, const char* loc, const int32_t* interf, const double* nx, const double* ny, int gh, int im,
                      int jm);
int bc_bc_symmetry_2d_d(double* w, double* wd, const char* loc, const int32_t* interf, const double* nx, const double* ny,
                        int gh, int im, int jm);
/* antisymmetry plane and pressure outlet (card_bl2d_fv_cgns.py:102 routineout = 'bc_pressure_2d'):
 * srcfv/borders/bc_antisymmetry.F90:1-79 + srcfv/tangent/bc_antisymmetry_d.f90; srcfv/borders/bc_pressure.F90:1-149 +
 * srcfv/tangent/bc_pressure_d.f90 (noref: Fortran logical as int; em must be 5) */
int bc_bc_antisymmetry_2d(double* w, const char* loc, const int32_t* interf, const double* nx, const double* ny, int gh, int im,
                          int jm);
int bc_bc_antisymmetry_2d_d(double* w, double* wd, const char* loc, const int32_t* interf, const double* nx, const double* ny,
                            int gh, int im, int jm);
int bc_bc_pressure_2d(double* w, const char* loc, const int32_t* interf, double pext, int noref, double gam, const double* nx,
                      const double* ny, int im, int jm, int gh, int em);
int bc_bc_pressure_2d_d(double* w, double* wd, const char* loc, const int32_t* interf, double pext, int noref, double gam,
                        const double* nx, const double* ny, int im, int jm, int gh, int em);
/* profile walls of the sensitivity driver (BROADCAST_npz_sens.py:1763 flinwall(w, wd, velprof, velprofd, 'Jlo', gam, interf, gh, im, jm);
 * card_bl2d_fv_npz_sens.py:105): srcfv/borders/bc_wall_blow_profile.F90:1-97 + srcfv/tangent/bc_wall_blow_profile_d.f90,
 * srcfv/borders/bc_wall_viscous_iso_profile.F90:1-107 + srcfv/tangent/bc_wall_viscous_iso_profile_d.f90.  The shipped tangents treat
 * the profile, gam and rgaz as active: velprofd / twallprofd (lm values), gamd, rgazd are their tangents. */
int bc_bc_wall_blow_profile_2d(double* w, const double* velprof, const char* loc, double gam, const int32_t* interf, int gh, int im,
                               int jm, int lm);
int bc_bc_wall_blow_profile_2d_d(double* w, double* wd, const double* velprof, const double* velprofd, const char* loc, double gam,
                                 double gamd, const int32_t* interf, int gh, int im, int jm, int lm);
int bc_bc_wall_viscous_iso_profile_2d(double* w, const double* twallprof, const char* loc, double gam, double rgaz,
                                      const int32_t* interf, int gh, int im, int jm, int lm);
int bc_bc_wall_viscous_iso_profile_2d_d(double* w, double* wd, const double* twallprof, const double* twallprofd, const char* loc,
                                        double gam, double gamd, double rgaz, double rgazd, const int32_t* interf, int gh, int im,
                                        int jm, int lm);
/* ---- Dirichlet fill from a table: srcfv/borders/bc_general.F90:6-31 (f_bnd.bc_general_2d: ghost layer de of line cell l takes
 *      field(l, de, :); the cards name it as the alternative inlet routine, card_bl2d_fv.py:107) and srcfv/tangent/bc_general_d.f90
 *      (f_lin.bc_general_2d_d: ghost tangents zeroed, w untouched).  field: (lm, gh1, em) Fortran order; em = 5, gh1 = gh. */
int bc_bc_general_2d(double* w, const char* loc, const int32_t* interf, const double* field, int gh, int im, int jm, int lm, int em,
                     int gh1);
int bc_bc_general_2d_d(double* w, double* wd, const char* loc, const int32_t* interf, const double* field, int gh, int im, int jm,
                       int lm);

/* srcfv/borders/jn_match.F90:3-66 (3-D arrays, em planes) and jn_match_geom.F90:7-69 (2-D arrays) */
int bc_jn_match_2d(double* wr, const int32_t* prr, int gh1r, int gh2r, int gh3r, int gh4r, int imr, int jmr,
                   const double* wd, const int32_t* prd, int gh1d, int gh2d, int gh3d, int gh4d, int imd, int jmd,
                   const int32_t* tr, int em);
int bc_jn_match_geom_2d(double* wr, const int32_t* prr, int gh1r, int gh2r, int gh3r, int gh4r, int imr, int jmr,
                        const double* wd, const int32_t* prd, int gh1d, int gh2d, int gh3d, int gh4d, int imd, int jmd,
                        const int32_t* tr);

/* ---- geometry: srcfv/geom/computegeom.F90:3-104 (f_geom.computegeom_2d), all arrays in place */
int bc_computegeom_2d(double* x0, double* y0, double* nx, double* ny, double* xc, double* yc, double* vol,
                      double* volf, int im, int jm, int gh);

/* ---- colouring seeds and COO scatter: misc/ComputeJacobian.f90 (f_misc.*); m,l,k are 0-based.
 *      jac/ia/ja have nbentry = 25*(2gh+1)^2*im*jm entries; one call writes the contiguous slot range of
 *      colour (m,l,k) (:524). */
int bc_testvector(double* wd, int m, int l, int k, int gh, int im, int jm);
int bc_testvector_partial(double* wd, int m, int l, int k, int gh, int im, int jm, int istart, int iend, int jstart,
                          int jend);
int bc_computejacobianfromjv(double* jac, int32_t* ia, int32_t* ja, const double* resd, int m, int l, int k, int gh,
                             int im, int jm, int64_t nbentry);
int bc_computejacobianfromjv_relaxed(double* jac, int32_t* ia, int32_t* ja, const double* resd, int m, int l, int k,
                                     int gh, int im, int jm, int64_t nbentry, const double* coefdiag);
int bc_computejacobianfromjv_relaxed_withjn(double* jac, int32_t* ia, int32_t* ja, const double* resd, int m, int l,
                                            int k, int gh, int im, int jm, int64_t nbentry, const double* coefdiag);
/* misc/ComputeJacobian.f90:1095-1204 -- two zones joined in i (cylinder.py:1159): zone n (0 or 1) uses the slot range shifted
 * by n * 25 (2gh+1)^2 im jm, writes only slots whose |jac| < mini and clears the others where the reference does; jac, ia,
 * ja are read AND written. */
int bc_computejacobianfromjv_relaxed_withjnandcheck(double* jac, int32_t* ia, int32_t* ja, const double* resd, int m, int l,
                                                    int k, int gh, int im, int jm, int64_t nbentry, const double* coefdiag,
                                                    double mini, int n);
int bc_computejacobianfromjv_withjn(double* jac, int32_t* ia, int32_t* ja, const double* resd, int m, int l, int k,
                                    int gh, int im, int jm, int64_t nbentry);
int bc_computejacobianfromdz(double* jac, int32_t* ia, int32_t* ja, const double* dz, int m, int l, int k, int gh,
                             int im, int jm, int64_t nbentry);

/* ---- spanwise operator rows: srcfv/dz/coeffs_5p_dz.F90:10-174, srcfv/dz/coeffs_5p_dz2.F90 (f_dz.coeffs_5p_dz,
 *      f_dz.coeffs_5p_dz2; BROADCAST_npz.py:1242-1243).  dz_out interior cells are written, its ghosts untouched. */
int bc_coeffs_5p_dz(double* dz_out, const double* w, const double* wd, const double* x0, const double* y0, const double* nx,
                    const double* ny, const double* xc, const double* yc, const double* vol, const double* volf, int gh, double cp,
                    double cv, double prandtl, double gam, double rgaz, double cs, double muref, double tref, double s_suth, int im,
                    int jm);
int bc_coeffs_5p_dz2(double* dz_out, const double* w, const double* wd, const double* x0, const double* y0, const double* nx,
                     const double* ny, const double* xc, const double* yc, const double* vol, const double* volf, int gh, double cp,
                     double cv, double prandtl, double gam, double rgaz, double cs, double muref, double tref, double s_suth, int im,
                     int jm);

/* ---- tangent of the spanwise operator rows w.r.t. the base flow: srcfv/tangentdz/coeffs_5p_dz_d.f90 (COEFFS_5P_DZ_D),
 *      srcfv/tangentdz/coeffs_5p_dz2_d.f90 (COEFFS_5P_DZ2_D) = f_lindz.coeffs_5p_dz_d / coeffs_5p_dz2_d of the sensitivity driver
 *      (BROADCAST_npz_sens.py:1768-1797, 2157-2185).  wd0 = variation of the base flow, wd = the mode.  As in the reference, dz_out
 *      is never assigned and the WHOLE of dz_outd is written (interior rows, ghost frame = 0). */
int bc_coeffs_5p_dz_d(double* dz_out, double* dz_outd, const double* w, const double* wd0, const double* wd, const double* x0,
                      const double* y0, const double* nx, const double* ny, const double* xc, const double* yc, const double* vol,
                      const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                      double muref, double tref, double s_suth, int im, int jm);
int bc_coeffs_5p_dz2_d(double* dz2_out, double* dz2_outd, const double* w, const double* wd0, const double* wd, const double* x0,
                       const double* y0, const double* nx, const double* ny, const double* xc, const double* yc, const double* vol,
                       const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                       double muref, double tref, double s_suth, int im, int jm);

/* ---- boundary tables from the initial field: set_bnd.f90:2-24 (f_init.set_bndbl_2d).  A host-side copy
 *      (field(j,depth,:) = w(1-depth,j,:), wbd(i,:) = w(i-gh,jm,:)); set-up, not on the device path. */
int bc_set_bndbl_2d(const double* w, double* field, double* wbd, int im, int jm, int gh);

/* ---- norms: srcfv/norm.F90:2-77 (f_norm.compute_norml2 / compute_norml2inf).  Reduction order on the
 *      device is a tree (the Fortran sums j-outer, i-inner sequentially). */
int bc_compute_norml2(double* norm, double* nmoy, const double* rhs, int im, int jm, int gh);
int bc_compute_norml2inf(double* norm, double* ninf, const double* rhs, int im, int jm, int gh);

/* =====================================================================================================
 * Device-pointer (resident) API.  Same semantics, no host transfers, asynchronous on `stream`.
 * `ndir` = 0 primal; 1 or 5 = number of tangent directions held in wd / residud as [ndir][5] planes.
 * ===================================================================================================== */
/* i-slab context of the calling thread (multi-GPU).  Between bcd_slab_begin and bcd_slab_end the device
 * entry points treat their (im, jm) block as columns ioff+1 .. ioff+im of a block of im_global columns: colouring seeds,
 * row/column numbers and the nearest-seed rules of the scatter use GLOBAL column indices; `edges` bit 0 / bit 1 says the
 * Ilo / Ihi side is a slab-internal edge whose gh halo columns hold the neighbour's cells (filled by the caller's halo
 * exchange), so velocity gradients are computed there instead of extrapolated (rhs/gradveloingh.F applies to physical
 * sides only) and the regular-row block kernels run up to that edge.  Arrays, rect arguments and slot layouts stay local. */
int bcd_slab_begin(int ioff, int im_global, int edges);
int bcd_slab_end(void);
/* Colour sharding (grids too small for i-slabs and the i-periodic O-mesh): the device colour loops called by this
 * thread visit only the (l,k) passes with c0 <= l*(2gh+1)+k < c1 (49 passes in all for gh = 3); c1 <= c0 = all colours.  Every
 * rank holds the whole state; the COO slots of a colour belong to exactly one rank, the host merges the filtered lists. */
int bcd_colour_range(int c0, int c1);
/* pitched copy of `height` rows of `width` BYTES between a host array and a device array (an i-slab of a Fortran-ordered
 * block is a set of pitched rows); kind 1 = host to device, 2 = device to host; asynchronous on `stream`. */
int bcd_memcpy2d(void* dst, long long dpitch, const void* src, long long spitch, long long width, long long height, int kind,
                 void* stream);
/* The residual in two parts for overlap with the halo exchange / boundary fills: part 1 = the tiles of the fused kernel that read
 * neither ghost cells nor slab halo columns (everything but the outermost ring of 32 x 9 tiles), part 2 = that ring, part 0 = all.
 * Part 1 may run while the ghosts are still being written; part 2 must follow the fills.  1 then 2 == 0, bit for bit. */
int bcd_residual_part(double* residu, const double* w, const double* nx, const double* ny, const double* vol,
                      const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                      double muref, double tref, double s_suth, double k2, double k4, int im, int jm, int wall, int part,
                      void* stream);
/* use_generic selects the kernel: 0 = default fused tile kernel (k_residual_fast), 1 = reference-shaped four-kernel
 * pipeline, 2 = fused kernel with persistent CTAs and TMA staging of w, 3 = first-generation fused kernel. All variants
 * compute the same residual (tests: 1e-12 of the plane maximum). */
int bcd_residual(double* residu, const double* w, const double* nx, const double* ny, const double* vol,
                 const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                 double muref, double tref, double s_suth, double k2, double k4, int im, int jm, int wall,
                 int use_generic, void* stream);
int bcd_tangent(double* residud, const double* w, const double* wd, int ndir, const double* nx, const double* ny,
                const double* vol, const double* volf, int gh, double cp, double cv, double prandtl, double gam,
                double rgaz, double cs, double muref, double tref, double s_suth, double k2, double k4, int im, int jm,
                int wall, const int32_t* rect /* null or {i0,i1,j0,j1} */, void* stream);
int bcd_bc_wall_viscous_adia(double* w, double* wd, int ndir, const char* loc, double gam, const int32_t* interf,
                             int gh, int im, int jm, void* stream);
int bcd_bc_no_reflexion(double* w, double* wd, int ndir, const double* wbd, const char* loc, const int32_t* interf,
                        const double* nx, const double* ny, double gam, int gh, int im, int jm, int lm, void* stream);
int bcd_bc_supandsubinlet(double* w, double* wd, int ndir, const char* loc, const int32_t* interf,
                          const double* field, const double* nx, const double* ny, double gam, int im, int jm, int lm,
                          int gh, void* stream);
int bcd_bc_extrapolate_o2(double* w, double* wd, int ndir, const char* loc, const int32_t* interf, int im, int jm,
                          int gh, void* stream);
int bcd_bc_general(double* w, double* wd, int ndir, const char* loc, const int32_t* interf, const double* field, int gh, int im,
                   int jm, int lm, void* stream);
int bcd_bc_wall_viscous_iso(double* w, double* wd, int ndir, double twall, const char* loc, double gam, double rgaz,
                            const int32_t* interf, int gh, int im, int jm, void* stream);
int bcd_bc_symmetry(double* w, double* wd, int ndir, const char* loc, const int32_t* interf, const double* nx,
                    const double* ny, int gh, int im, int jm, int anti /* 1 = bc_antisymmetry_2d */, void* stream);
/* blow = 1: bc_wall_blow_profile_2d, 0: bc_wall_viscous_iso_profile_2d; prof / profd: device arrays of lm values (profd may be NULL) */
int bcd_bc_wall_profile(double* w, double* wd, int ndir, int blow, const double* prof, const double* profd, const char* loc,
                        double gam, double gamd, double rgaz, double rgazd, const int32_t* interf, int gh, int im, int jm, int lm,
                        void* stream);
int bcd_bc_pressure(double* w, double* wd, int ndir, const char* loc, const int32_t* interf, double pext, int noref,
                    double gam, const double* nx, const double* ny, int im, int jm, int gh, void* stream);
int bcd_jn_match(double* wr, const int32_t* prr, int gh1r, int gh2r, int gh3r, int gh4r, int imr, int jmr,
                 const double* wd, const int32_t* prd, int gh1d, int gh2d, int gh3d, int gh4d, int imd, int jmd,
                 const int32_t* tr, int em, void* stream);
int bcd_testvector(double* wd, int ndir, int m, int l, int k, int gh, int im, int jm, const int32_t* zone,
                   void* stream);
/* kind: 0 jv, 1 jv_relaxed, 2 dz, 3 jv_relaxed_withjn, 4 jv_withjn, 5 jv_dbyvol, 6 jv_relaxed_dbyvol */
int bcd_scatter(int kind, double* seg_jac, int32_t* seg_ia, int32_t* seg_ja, const double* resd, int m, int l, int k,
                int gh, int im, int jm, const double* coefdiag, const double* vol, void* stream);
/* spanwise operator rows on the device: which = 1 (d/dz) or 2 (d2/dz2); ndir = 1 or 5 tangent directions in wd / dz_out */
int bcd_dz(double* dz_out, const double* w, const double* wd, int ndir, int which, const double* nx, const double* ny,
           const double* vol, const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
           double muref, double tref, double s_suth, int im, int jm, const int32_t* rect, void* stream);
/* tangent of both operator rows w.r.t. the base flow in ONE pass over device arrays (either output may be NULL); interior cells of
 * `rect` (NULL = all) are written, nothing else */
int bcd_dz_tangent(double* dz_outd, double* dz2_outd, const double* w, const double* wd0, const double* wd, const double* nx,
                   const double* ny, const double* vol, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                   double muref, double tref, double s_suth, int im, int jm, const int32_t* rect, void* stream);
/* out10 (device): sum r^2 per equation [5], sum r^10 per equation [5] */
int bcd_norm_sums(double* out10, const double* rhs, int im, int jm, int gh, void* stream);

/* =====================================================================================================
 * Whole colour loop on the device (resident mode).
 *
 * bc_desc_t describes one boundary operation of the driver's ordered list (BROADCAST_npz.py:1018-1021,
 * 1079-1082; handleBC.py:129-242).  Tables are DEVICE pointers.
 * ===================================================================================================== */
#define BC_KIND_INLET 1   /* bc_supandsubinlet_2d   table = field(lm,gh,5) */
#define BC_KIND_NOREF 2   /* bc_no_reflexion_2d     table = wbd(lm,5)      */
#define BC_KIND_EXTRAP 3  /* bc_extrapolate_o2_2d                          */
#define BC_KIND_WALL 4    /* bc_wall_viscous_adia_2d                       */
#define BC_KIND_JOIN 5    /* jn_match_2d of the block onto itself: window = receiver, prd = donor */
#define BC_KIND_WALL_ISO 6      /* bc_wall_viscous_iso_2d   param = {twall, rgaz}          */
#define BC_KIND_SYMMETRY 7      /* bc_symmetry_2d                                          */
#define BC_KIND_ANTISYMMETRY 8  /* bc_antisymmetry_2d                                      */
#define BC_KIND_PRESSURE 9      /* bc_pressure_2d           param = {pext, noref (0 / 1)}  */
#define BC_KIND_WALL_BLOW_PROFILE 10  /* bc_wall_blow_profile_2d        table = velprof(lm)   (passive in the colour loops) */
#define BC_KIND_WALL_ISO_PROFILE 11   /* bc_wall_viscous_iso_profile_2d table = twallprof(lm), param = {unused, rgaz}       */
typedef struct {
  int32_t kind;
  char loc[4];
  int32_t window[4]; /* interf (kinds 1-4, 6-9) or prr (kind 5): imin, jmin, imax, jmax */
  int32_t prd[4];    /* kind 5 only */
  int32_t tr[2];     /* kind 5 only */
  int32_t lm;
  const double* table;
  double param[2];   /* scalar arguments of kinds 6, 9 and 11 */
} bc_desc_t;

/* Reference colour loop (BROADCAST_npz.py:1068-1127 / cylinder.py:941-978) entirely on the device:
 * for every colour (l,k): seeds for the 5 variables at once (vector tangent mode), linearised boundary
 * fills in list order, tangent of the residual, scatter.  Output = the reference's own COO arrays
 * (jac, ia, ja of length 25*(2gh+1)^2*im*jm, slot order of misc/ComputeJacobian.f90:524), on the device.
 * scatter_kind as in bcd_scatter.  The primal list is applied to w once before the first colour (handleBC.applyBC(mode 1) of the
 * cylinder driver re-applies it in every colour: same ghosts, they depend on interior cells only).  rect (or null) restricts the rows
 * that are evaluated to cells i0..i1 x j0..j1 (others keep their previous content); with compact != 0 the
 * outputs have 25*(2gh+1)^2*wi*wj entries, slot order as above over the rectangle's own (wi x wj) index space
 * (ia/ja stay global). */
int bcd_jacobian_coo(double* jac, int32_t* ia, int32_t* ja, double* w, const double* nx, const double* ny,
                     const double* vol, const double* volf, int gh, double cp, double cv, double prandtl, double gam,
                     double rgaz, double cs, double muref, double tref, double s_suth, double k2, double k4, int im,
                     int jm, int wall, const bc_desc_t* bcs, int nbcs, int scatter_kind, const double* coefdiag,
                     const int32_t* rect, int compact, void* stream);
/* The same colour loop restricted to the rows of up to four rectangles (the boundary strips of the hybrid assembly),
 * all rectangles in the same 49 passes; output: one compact COO triple per rectangle (as bcd_jacobian_coo with compact). */
int bcd_jacobian_strips(int nrect, const int32_t* rects /* [nrect][4] = i0,i1,j0,j1 */, double* const* jac, int32_t* const* ia,
                        int32_t* const* ja, double* w, const double* nx, const double* ny, const double* vol, const double* volf,
                        int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs, double muref, double tref,
                        double s_suth, double k2, double k4, int im, int jm, int wall, const bc_desc_t* bcs, int nbcs,
                        int scatter_kind, const double* coefdiag, void* stream);
/* Colour loop of the spanwise operators (BROADCAST_npz.py:1231-1246) on the device: COO triplets of Dz (jac1, ia1, ja1)
 * and Dz2 (jac2, ia2, ja2) in the reference's slot order, scatter rule computejacobianfromdz (misc/ComputeJacobian.f90:708-779).
 * Either triplet may be null. */
int bcd_dz_coo(double* jac1, int32_t* ia1, int32_t* ja1, double* jac2, int32_t* ia2, int32_t* ja2, double* w, const double* nx,
               const double* ny, const double* vol, const double* volf, int gh, double cp, double cv, double prandtl, double gam,
               double rgaz, double cs, double muref, double tref, double s_suth, int im, int jm, const bc_desc_t* bcs, int nbcs,
               void* stream);
/* Colour loop of the sensitivity driver (BROADCAST_npz_sens.py:1741-1800) on the device: d/dw [Dz(w) mode] and d/dw [Dz2(w) mode], column by
 * column (seeds, linearised boundary fills of the list, f_lindz.coeffs_5p_dz_d / coeffs_5p_dz2_d with wd0 = seed and wd = mode,
 * computejacobianfromdz).  wmoder / wmodei: real and imaginary part of the mode (wmodei may be NULL); jac1r, jac1i share ia1 / ja1 and
 * jac2r, jac2i share ia2 / ja2 (the driver's IAdz, JAdz, IAdz2, JAdz2); lists of 25 (2gh+1)^2 im jm entries in the reference's slot
 * order; any value list may be NULL. */
int bcd_dz_tangent_coo(double* jac1r, double* jac1i, int32_t* ia1, int32_t* ja1, double* jac2r, double* jac2i, int32_t* ia2,
                       int32_t* ja2, double* w, const double* wmoder, const double* wmodei, const double* nx, const double* ny,
                       const double* vol, int gh, double cp, double cv, double prandtl, double gam, double rgaz, double cs,
                       double muref, double tref, double s_suth, int im, int jm, const bc_desc_t* bcs, int nbcs, void* stream);
/* Direct block-Jacobian of the regular interior rows (rows whose stencil touches neither ghost cells nor
 * the wall rows: gh+1 <= i <= im-gh, gh+1 <= j <= jm-gh) into the fixed 29-slot pattern, without colouring:
 * values[slot][e][m][cell], cell = (i-1) + (j-1)*im, slot s <-> column-cell offset given by
 * bcd_jacobian_slots().  Entry (e,m) of a block is jac = -d residu(i,j,e) / d w(i+di,j+dj,m) (+ coefdiag(i,j)
 * on the diagonal if coefdiag != null): what the reference's colour loop + computejacobianfromjv_relaxed
 * attribute to that pair (misc/ComputeJacobian.f90:518-569).  rect (or null) restricts the rows. */
int bcd_jacobian_slots(int32_t* offsets /* [29][2] (di,dj) */);
int bcd_jacobian_interior(double* values, const double* w, const double* nx, const double* ny, const double* vol,
                          const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz,
                          double cs, double muref, double tref, double s_suth, double k2, double k4, int im, int jm,
                          const double* coefdiag, const int32_t* rect, void* stream);
/* Same result by the direct forward-AD kernels (one launch per column offset, every other cell passive at
 * compile time): the slower cross-check of the face-linearisation path used by bcd_jacobian_interior. */
int bcd_jacobian_interior_ad(double* values, const double* w, const double* nx, const double* ny, const double* vol,
                             const double* volf, int gh, double cp, double cv, double prandtl, double gam, double rgaz,
                             double cs, double muref, double tref, double s_suth, double k2, double k4, int im, int jm,
                             const double* coefdiag, const int32_t* rect, void* stream);
/* Hybrid block Jacobian -> CSR row block on the device (what the reference does on the host: remove_zero_jac, BROADCAST_npz.py:129-135;
 * csr_matrix((Jac,(IA,JA))), misc/PETSc_func.py:85; the division by the cell volume, BROADCAST_npz.py:1206-1209).
 * values = the 29-block array of bcd_jacobian_interior (region = i0,i1,j0,j1 of its valid rows), sjac/sia/sja/slen = the compact COO
 * lists of bcd_jacobian_strips, srect their rectangles.  Local row = e + 5 (j-1) + 5 jm (i-1); columns global, ascending in a row.
 * Step 1 (…_indptr): kept entries per row (|v| > thresh) and their exclusive scan; indptr has 5 im jm + 1 int64 entries, the last
 * one is nnz.  counts: 5 im jm + 1 ints, bsum: 5 im jm / 2048 + 2 int64 of device work space.
 * Step 2 (…_fill): indices / data of nnz entries; vol != null divides each row by the volume of its cell; cursor: 5 im jm + 1 ints. */
int bcd_hybrid_csr_indptr(long long* indptr, int32_t* counts, long long* bsum, const double* values, const int32_t* region,
                          int nstrip, const double* const* sjac, const int32_t* const* sia, const long long* slen, double thresh,
                          int gh, int im, int jm, void* stream);
/* Counting fused into the assembly: bcd_jacobian_interior that also leaves, in counts[0 .. 5 im jm] (zeroed first), the number of
 * entries |v| > thresh of every regular row, taken while the block values are in registers; bcd_hybrid_csr_indptr_counted then only
 * adds the strip rows and scans (one pass over the block values less: 5.8 KB per cell). */
int bcd_jacobian_interior_counted(double* values, int32_t* counts, double thresh, const double* w, const double* nx,
                                  const double* ny, const double* vol, const double* volf, int gh, double cp, double cv,
                                  double prandtl, double gam, double rgaz, double cs, double muref, double tref, double s_suth,
                                  double k2, double k4, int im, int jm, const double* coefdiag, const int32_t* rect, void* stream);
int bcd_hybrid_csr_indptr_counted(long long* indptr, int32_t* counts, long long* bsum, const int32_t* region, int nstrip,
                                  const double* const* sjac, const int32_t* const* sia, const long long* slen, double thresh,
                                  int gh, int im, int jm, void* stream);
int bcd_hybrid_csr_fill(int32_t* indices, double* data, int32_t* cursor, const long long* indptr, const double* values,
                        const int32_t* region, int nstrip, const int32_t* srect, const double* const* sjac,
                        const int32_t* const* sia, const int32_t* const* sja, const long long* slen, double thresh,
                        const double* vol, int gh, int im, int jm, void* stream);
/* ---- adjoint operator (first step): CSR of A^T from the CSR row block of A on the device.  The adjoint-baseflow
 *      and optimal-forcing drivers transpose the assembled Jacobian (cylinder.py:1090-1177; misc/PETSc_func.py createTranspose).
 *      Step 1: tptr[0 .. ncols] (int64); the caller reads tptr[ncols] = nnz and allocates tind (int32) / tdat.  Step 2: fill, the
 *      entries of every transposed row in ascending column order (= scipy's csr_matrix(A.T) on a matrix with sorted indices).
 *      counts / cursor: ncols + 1 int32 of work space; bsum: ncols / 2048 + 2 int64.  row0: first global row of the block. */
int bcd_csr_transpose_indptr(long long* tptr, int32_t* counts, long long* bsum, const int32_t* indices, long long nnz,
                             long long ncols, void* stream);
int bcd_csr_transpose_fill(int32_t* tind, double* tdat, int32_t* cursor, const long long* tptr, const long long* indptr,
                           const int32_t* indices, const double* data, long long nrows, long long row0, long long ncols,
                           void* stream);

/* ---- the step after the assembly: Newton correction A dw = res on the device.  The reference hands the CSR to
 *      PETSc / MUMPS LU on the host (misc/PETSc_func.py:137-152 kspLUPetsc, :247-263 iterNewton: dw = A^-1 res, then
 *      w += dw, BROADCAST_npz.py:1157-1172) and leaves its `lasolver == 'gmres'` branch unimplemented (:1043-1047).  Here: restarted
 *      GMRES, right-preconditioned by the inverse 5 x 5 diagonal blocks, everything resident (csrc/solve.cu).  The adjoint systems
 *      (cylinder.py:1090-1177) run the same entry points on the transposed CSR (bcd_csr_transpose_*).  One device, whole matrix. */
/* y = A x for a CSR row block (indptr int64, global int32 columns) */
int bcd_csr_spmv(double* y, const long long* indptr, const int32_t* indices, const double* data, const double* x,
                 long long nrows, void* stream);
/* dinv[25][ncell]: inverse of the diagonal 5 x 5 block of every cell (rows 5 c .. 5 c + 4 of the block, columns col0 + 5 c ...);
 * *nbad (device int) = number of singular blocks, replaced by the identity */
int bcd_block_jacobi_setup(double* dinv, int32_t* nbad, const long long* indptr, const int32_t* indices, const double* data,
                           long long ncell, long long col0, void* stream);
int bcd_block_jacobi_apply(double* z, const double* dinv, const double* r, long long ncell, void* stream);
/* doubles of device work space bcd_gmres needs for n unknowns and the given restart length (1 <= restart < 64); -1 if invalid */
long long bcd_gmres_work_doubles(long long n, int restart);
/* x: start vector in, solution out; dinv: bcd_block_jacobi_setup output or null; side = 1: preconditioner on the left (the residual
 * the iteration controls is M^-1 (b - A x): free of the row scaling of the finite-volume Jacobian), 0: on the right; stops at a
 * relative (preconditioned) residual of rtol or after maxit products; info[0] = matrix-vector products, info[1] = 1 if converged (host
 * ints); relres[0] = final TRUE relative residual ||b - A x|| / ||b||, relres[1] = the controlled one (host doubles) */
int bcd_gmres(double* x, const double* b, const long long* indptr, const int32_t* indices, const double* data,
              const double* dinv, long long n, int restart, int maxit, double rtol, int side, double* work,
              long long work_len, int32_t* info, double* relres, void* stream);
/* isothermal-wall scheme variant in resident mode: the wall flux of the calling thread's next bcd_* launches (residual, tangent,
 * colour loops, strips) is rhs/fluxwall_iso.F with this twall while on != 0 (bc_flux_num_dnc5_iso_2d sets it around its own call) */
int bcd_wall_iso(int on, double twall);
/* primal boundary fill of a whole list */
int bcd_apply_bcs(double* w, const double* nx, const double* ny, double gam, int gh, int im, int jm,
                  const bc_desc_t* bcs, int nbcs, void* stream);

/* ---- halo exchange of the i-slabs by peer stores over NVLink (csrc/halo.cu; the periodic cut of an O-mesh is the
 *      same exchange between the first and the last slab, cylinder.py:499-527).  One process per GPU: every rank creates a mailbox
 *      (`rows` = planes * (jm + 2 gh) rows of gh doubles per side), hands its 64-byte IPC handle to its neighbours (any host
 *      transport: the Python layer uses torch.distributed.all_gather_object) and connects theirs: side 0 = left neighbour
 *      (fills my Ilo halo), side 1 = right neighbour.  bcd_halo_exchange = two launches on `stream` (push into the neighbours'
 *      mailboxes + release flag; wait for my own flags + unpack into the halo columns of w); step numbers live on the device, so
 *      the pair can be captured in a CUDA graph.  bcd_halo_error: 1 after an unpack that saw no delivery for ~2 s. */
int bcd_halo_create(void** handle, int gh, long long rows, unsigned char* ipc_handle_out64);
int bcd_halo_connect(void* handle, int side, const unsigned char* ipc_handle64, void* same_process_mailbox, int peer_device);
void* bcd_halo_mailbox(void* handle);
void* bcd_halo_peer(void* handle, int side);
int bcd_halo_exchange(void* handle, double* w, long long rows, int ni, int gh, void* stream);
int bcd_halo_error(void* handle);
int bcd_halo_destroy(void* handle);
/* ---- capture of a sequence of bcd_* calls issued on `stream` between begin and end into a CUDA graph (thread-local capture mode),
 *      replayed by bcd_graph_launch: the per-step sequence exchange + boundary fills + residual as ONE host call.  Every entry
 *      point used inside must have run once before (scratch allocations and kernel attributes are not capturable). */
int bcd_graph_begin(void* stream);
int bcd_graph_end(void* stream, void** exec_out);
int bcd_graph_launch(void* exec, void* stream);
int bcd_graph_destroy(void* exec);

/* =====================================================================================================
 * Resident-mode context for hosts without a device-memory library of their own (C, Fortran via ISO_C_BINDING; INTEGRATION.md
 * section 3): one block stays on the device between calls.  All array arguments are HOST arrays in the layout of the bc_* entry
 * points (bc_desc_t.table included: field(lm,gh,5) / wbd(lm,5) Fortran arrays, copied at set time).  Sequence of the reference
 * drivers: create -> set_geometry (after f_geom.computegeom_2d) -> set_bcs -> upload_state -> residual [-> norms / download_residual]
 * (BROADCAST_npz.py:1018-1035) -> jacobian_csr -> download_csr (BROADCAST_npz.py:1068-1127, 129-135, 1206-1209;
 * misc/PETSc_func.py:71-95).  Every number comes from the bcd_* entry points above.
 * ===================================================================================================== */
typedef struct bcast_ctx bcast_ctx_t;
int bcast_ctx_create(bcast_ctx_t** ctx, int im, int jm, int gh, double cp, double cv, double prandtl, double gam, double rgaz,
                     double cs, double muref, double tref, double s_suth, double k2, double k4, int wall);
int bcast_ctx_destroy(bcast_ctx_t* ctx);
int bcast_ctx_set_geometry(bcast_ctx_t* ctx, const double* nx, const double* ny, const double* vol, const double* volf);
int bcast_ctx_set_bcs(bcast_ctx_t* ctx, const bc_desc_t* bcs, int nbcs);
int bcast_ctx_upload_state(bcast_ctx_t* ctx, const double* w);
int bcast_ctx_download_state(bcast_ctx_t* ctx, double* w);
int bcast_ctx_apply_bcs(bcast_ctx_t* ctx);
int bcast_ctx_residual(bcast_ctx_t* ctx);
int bcast_ctx_download_residual(bcast_ctx_t* ctx, double* res);
int bcast_ctx_norms(bcast_ctx_t* ctx, double* norm5, double* ninf5);
/* coefdiag: host (im,jm) array or NULL; scatter_kind 1 = jv_relaxed, 3 = jv_relaxed_withjn, -1 = by the boundary list */
int bcast_ctx_jacobian_csr(bcast_ctx_t* ctx, const double* coefdiag, int divide_by_vol, double thresh, int scatter_kind,
                           long long* nnz);
int bcast_ctx_download_csr(bcast_ctx_t* ctx, long long* indptr, int32_t* indices, double* data);

#ifdef __cplusplus
}
#endif
#endif /* BROADCAST_B200_H */
