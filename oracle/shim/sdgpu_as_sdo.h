/*
 * oracle/shim/sdgpu_as_sdo.h  --  TEST INFRASTRUCTURE ONLY.
 * Lets the patched reference host (integration/twoSD_src.patch + integration/sdgpu_hooks.c) link against the restated CPU oracle
 * (oracle/libsdoracle.so, prefix sdo_) instead of libsdgpu.so, so that the host patch also runs in the CPU test suite, where no
 * GPU exists.  Force-included (-include) by the `hooks` target of oracle/Makefile for the CPU flavour only.
 */
#ifndef SDGPU_AS_SDO_H
#define SDGPU_AS_SDO_H
#define sdgpu_create                      sdo_create
#define sdgpu_reset                       sdo_reset
#define sdgpu_destroy                     sdo_destroy
#define sdgpu_last_error                  sdo_last_error
#define sdgpu_get_counts                  sdo_get_counts
#define sdgpu_calc_omega                  sdo_calc_omega
#define sdgpu_calc_delta                  sdo_calc_delta
#define sdgpu_update_dual                 sdo_update_dual
#define sdgpu_update_dual_col             sdo_update_dual_col
#define sdgpu_basis_find_or_append        sdo_basis_find_or_append
#define sdgpu_basis_set_obs_feasible_col  sdo_basis_set_obs_feasible_col
#define sdgpu_basis_set_obs_feasible_row  sdo_basis_set_obs_feasible_row
#define sdgpu_basis_set_feas_data         sdo_basis_set_feas_data
#define sdgpu_check_feasibility_basis     sdo_check_feasibility_basis
#define sdgpu_sd_cut                      sdo_sd_cut
#define sdgpu_dual_stability              sdo_dual_stability
#define sdgpu_cut_heights                 sdo_cut_heights
#define sdgpu_feas_cuts                   sdo_feas_cuts
#define sdgpu_reform_cut                  sdo_reform_cut
#endif
