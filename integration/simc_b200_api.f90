! simc_b200_api.f90 -- ISO_C_BINDING mirror of include/simc_b200.h (ABI version 2).
!
! Compile next to the reference's sources (free form, any -fdefault-real-8 setting: every kind below is explicit)
! and link the driver with -lsimc_b200.  The derived types repeat the C structures field by field, in the same
! order; tests/test_shim_layout.py parses this file and checks names, order and byte offsets against the C layout.
! The calls that replace the event loop of `program simc` (simc.f:169-351) are in simc_b200_shim.f.
module simc_b200_api
  use iso_c_binding
  implicit none

  integer(c_int32_t), parameter :: SIMC_B200_ABI_VERSION = 3
  integer(c_int), parameter :: SIMC_NHIST = 50, SIMC_H_PER_SET = 8, SIMC_NSTOP = 64
  integer(c_int), parameter :: SIMC_NTUPLE_MAXCOL = 68

  ! ---- basic records (modules.f:5-12, 35-38, 170-184, 202-215, 150-165, 195)
  type, bind(C) :: simc_cut
    real(c_double) :: min, max
  end type
  type, bind(C) :: simc_range
    real(c_double) :: lo, hi
  end type
  type, bind(C) :: simc_arm_cuts
    type(simc_cut) :: delta, yptar, xptar, z
  end type
  type, bind(C) :: simc_arm_limits
    type(simc_cut) :: delta, yptar, xptar, E
  end type
  type, bind(C) :: simc_edge_arm
    type(simc_cut) :: E, yptar, xptar
  end type
  type, bind(C) :: simc_edge
    type(simc_edge_arm) :: e, p
    type(simc_cut) :: Em, Pm, Mrec, Trec, Trec_struck
  end type
  type, bind(C) :: simc_gen_limits
    type(simc_arm_limits) :: e, p
    type(simc_cut) :: sumEgen, Trec
    real(c_double) :: xwid, ywid
  end type
  type, bind(C) :: simc_spectrometer
    real(c_double) :: P, theta, cos_th, sin_th, phi
    real(c_double) :: off_x, off_y, off_z, off_xptar, off_yptar
  end type
  type, bind(C) :: simc_axis
    real(c_double) :: min, bin
  end type
  ! target_info (target.inc:37-49), what the loop reads
  type, bind(C) :: simc_target
    real(c_double) :: A, Z, N, mass_amu, M, mrec_amu, Mrec, rho, thick, angle, abundancy
    real(c_double) :: length, zoffset, X0, X0_cm, L1, L2, fr1, fr2, xoffset, yoffset
    real(c_double) :: Coulomb_ave, Coulomb_min, Coulomb_max, Coulomb_constant
    real(c_double) :: Mtar_struck, Mrec_struck
    integer(c_int32_t) :: fr_pattern, can
  end type

  ! ---- run constants: /gnrl/ (simulate.inc:91-113), /radccom/ (radc.inc:13-19), /target_info/ (target.inc:52-53),
  !      /decd/ (simulate.inc:153-158), histogram axes
  type, bind(C) :: simc_run_config
    integer(c_int32_t) :: abi_version
    integer(c_int32_t) :: doing_phsp, doing_hyd_elast, doing_deuterium, doing_heavy, doing_eep
    integer(c_int32_t) :: doing_pion, doing_kaon, doing_delta, doing_rho, doing_semi
    integer(c_int32_t) :: doing_hydpi, doing_deutpi, doing_hepi
    integer(c_int32_t) :: doing_hydkaon, doing_deutkaon, doing_hekaon
    integer(c_int32_t) :: doing_hydsemi, doing_deutsemi
    integer(c_int32_t) :: doing_semipi, doing_semika
    integer(c_int32_t) :: do_fermi
    integer(c_int32_t) :: doing_hplus, doing_decay
    integer(c_int32_t) :: which_pion, which_kaon
    integer(c_int32_t) :: using_rad, using_Eloss, using_Coulomb, correct_Eloss, correct_raster
    integer(c_int32_t) :: mc_smear, hard_cuts
    integer(c_int32_t) :: using_E_arm_montecarlo, using_P_arm_montecarlo
    integer(c_int32_t) :: electron_arm, hadron_arm
    integer(c_int32_t) :: using_HMScoll, using_SHMScoll, use_benhar_sf
    integer(c_int32_t) :: rad_flag, extrad_flag, intcor_mode, use_expon, use_offshell_rad
    integer(c_int32_t) :: doing_tail(3)
    integer(c_int32_t) :: hardwired_rad
    integer(c_int32_t) :: deForest_flag
    integer(c_int32_t) :: doing_pizero, pizero_ngamma
    integer(c_int32_t) :: using_tgt_field, pad_flags
    real(c_double) :: Mh, Mh2, Ebeam, dEbeam, Ebeam_vertex_ave
    real(c_double) :: dE_edge_test, Egamma_gen_max, ctau, transparency
    real(c_double) :: drift_to_cal
    real(c_double) :: targ_Bangle, targ_Bphi, targ_pol, sign_hadron
    real(c_double) :: etatzai, Egamma_tot_max, Egamma1_max, Egamma2_max, Egamma3_max, Egamma_res_limit
    type(simc_gen_limits) :: gen
    type(simc_spectrometer) :: spec_e, spec_p
    type(simc_cut) :: cuts_Em, cuts_Pm
    type(simc_edge) :: edge, VERTEXedge
    type(simc_arm_cuts) :: SPedge_e, SPedge_p
    real(c_double) :: slop_MC_e_used(3), slop_MC_p_used(3)
    type(simc_target) :: targ
    type(simc_axis) :: hist_axis(8,3)
    real(c_double) :: w_ref
  end type

  ! 128-bit two's-complement fixed-point sum: value = (hi*2**64 + lo) * 2**qexp, lo unsigned
  type, bind(C) :: simc_fixed128
    integer(c_int64_t) :: lo
    integer(c_int64_t) :: hi
    integer(c_int32_t) :: qexp
    integer(c_int32_t) :: pad
  end type

  ! ---- what the loop leaves behind (simc.f:229-350); C arrays a[i][j][k] are Fortran a(k,j,i)
  type, bind(C) :: simc_accum
    integer(c_int64_t) :: ntried, nsuccess, ncontribute, npasscuts, ncontribute_no_rad_proton
    type(simc_fixed128) :: wtcontribute
    type(simc_fixed128) :: sum_sigcc
    type(simc_fixed128) :: sumerr(8), sumerr2(8)
    type(simc_fixed128) :: hist_w(50,6)
    integer(c_int64_t) :: hist_n(50,8,3)
    type(simc_range) :: contrib(32)
    type(simc_range) :: slop(8)
    integer(c_int64_t) :: stop(64,2)
    integer(c_int64_t) :: transp_calls(48,2)
    integer(c_int64_t) :: unsupported
    integer(c_int64_t) :: nonfinite
  end type

  type, bind(C) :: simc_results
    real(c_double) :: luminosity, genvol, normfac, yield, central_sigcc_ave
    integer(c_int64_t) :: nevent
    real(c_double) :: aveerr(8), resol(8)
  end type

  interface
    ! ---- lifecycle
    integer(c_int) function simc_b200_abi_version() bind(C, name='simc_b200_abi_version')
      import
    end function
    integer(c_int) function simc_b200_create(cfg, device, h) bind(C, name='simc_b200_create')
      import
      type(simc_run_config), intent(in) :: cfg
      integer(c_int), value :: device
      type(c_ptr), intent(out) :: h
    end function
    subroutine simc_b200_destroy(h) bind(C, name='simc_b200_destroy')
      import
      type(c_ptr), value :: h
    end subroutine
    type(c_ptr) function simc_b200_last_error(h) bind(C, name='simc_b200_last_error')
      import
      type(c_ptr), value :: h
    end function
    integer(c_int64_t) function simc_b200_sizeof(which) bind(C, name='simc_b200_sizeof')
      import
      integer(c_int), value :: which
    end function
    integer(c_int) function simc_b200_set_mode(h, strict_mode) bind(C, name='simc_b200_set_mode')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: strict_mode
    end function
    integer(c_int) function simc_b200_set_compiled_maps(h, on) bind(C, name='simc_b200_set_compiled_maps')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: on
    end function
    integer(c_int) function simc_b200_sync(h) bind(C, name='simc_b200_sync')
      import
      type(c_ptr), value :: h
    end function
    ! ---- tables: transp_init + mc_*_recon loaders, sf_lookup_init, theory_init, deut.dat, SetCtq5, fDSS, sigmaid
    integer(c_int) function simc_b200_load_optics(h, arm, fwd, rec) bind(C, name='simc_b200_load_optics')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: arm
      character(kind=c_char), intent(in) :: fwd(*), rec(*)
    end function
    integer(c_int) function simc_b200_load_sf_file(h, path, proton_flag) bind(C, name='simc_b200_load_sf_file')
      import
      type(c_ptr), value :: h
      character(kind=c_char), intent(in) :: path(*)
      integer(c_int), value :: proton_flag
    end function
    integer(c_int) function simc_b200_set_sf_table(h, n_pm, n_em, pm, em, sf) bind(C, name='simc_b200_set_sf_table')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: n_pm, n_em
      real(c_double), intent(in) :: pm(*), em(*), sf(*)
    end function
    integer(c_int) function simc_b200_set_sf_em_widths(h, n_em, dem) bind(C, name='simc_b200_set_sf_em_widths')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: n_em
      real(c_double), intent(in) :: dem(*)
    end function
    integer(c_int) function simc_b200_load_theory_file(h, path) bind(C, name='simc_b200_load_theory_file')
      import
      type(c_ptr), value :: h
      character(kind=c_char), intent(in) :: path(*)
    end function
    integer(c_int) function simc_b200_set_theory_table(h, n_shells, absorption, e_fermi, nprot, em, emsig, bs_norm, &
                                                       n_pm, pm_first, pm_bin, rho) bind(C, name='simc_b200_set_theory_table')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: n_shells
      real(c_double), value :: absorption, e_fermi
      real(c_double), intent(in) :: nprot(*), em(*), emsig(*), bs_norm(*), pm_first(*), pm_bin(*), rho(*)
      integer(c_int32_t), intent(in) :: n_pm(*)
    end function
    integer(c_int) function simc_b200_load_pfermi_file(h, path) bind(C, name='simc_b200_load_pfermi_file')
      import
      type(c_ptr), value :: h
      character(kind=c_char), intent(in) :: path(*)
    end function
    integer(c_int) function simc_b200_set_pfermi_table(h, n, pval, mprob) bind(C, name='simc_b200_set_pfermi_table')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: n
      real(c_double), intent(in) :: pval(*), mprob(*)
    end function
    integer(c_int) function simc_b200_load_cteq5_file(h, path) bind(C, name='simc_b200_load_cteq5_file')
      import
      type(c_ptr), value :: h
      character(kind=c_char), intent(in) :: path(*)
    end function
    integer(c_int) function simc_b200_load_fdss_file(h, path) bind(C, name='simc_b200_load_fdss_file')
      import
      type(c_ptr), value :: h
      character(kind=c_char), intent(in) :: path(*)
    end function
    integer(c_int) function simc_b200_load_maid_file(h, ipi, path) bind(C, name='simc_b200_load_maid_file')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: ipi
      character(kind=c_char), intent(in) :: path(*)
    end function
    ! field of the polarised target: replaces trgInit (trg_track.f:243-347, simc.f:154)
    integer(c_int) function simc_b200_load_field_file(h, path) bind(C, name='simc_b200_load_field_file')
      import
      type(c_ptr), value :: h
      character(kind=c_char), intent(in) :: path(*)
    end function
    integer(c_int) function simc_b200_set_field_map(h, bz, br) bind(C, name='simc_b200_set_field_map')
      import
      type(c_ptr), value :: h
      real(c_double), intent(in) :: bz(*), br(*)
    end function
    integer(c_int) function simc_b200_field_batch(h, spect, theta_deg, n, in_soa, out_soa) bind(C, name='simc_b200_field_batch')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: spect
      real(c_double), value :: theta_deg
      integer(c_int64_t), value :: n
      real(c_double), intent(in) :: in_soa(*)
      real(c_double), intent(out) :: out_soa(*)
    end function
    integer(c_int) function simc_b200_load_saghai_files(h, dir) bind(C, name='simc_b200_load_saghai_files')
      import
      type(c_ptr), value :: h
      character(kind=c_char), intent(in) :: dir(*)
    end function
    ! ---- the loop (simc.f:169-351)
    integer(c_int) function simc_b200_accum_clear(h, acc) bind(C, name='simc_b200_accum_clear')
      import
      type(c_ptr), value :: h
      type(simc_accum), intent(out) :: acc
    end function
    integer(c_int) function simc_b200_run(h, first_try, n_tries, seed, acc) bind(C, name='simc_b200_run')
      import
      type(c_ptr), value :: h
      integer(c_int64_t), value :: first_try, n_tries, seed
      type(simc_accum), intent(inout) :: acc
    end function
    integer(c_int) function simc_b200_run_async(h, first_try, n_tries, seed) bind(C, name='simc_b200_run_async')
      import
      type(c_ptr), value :: h
      integer(c_int64_t), value :: first_try, n_tries, seed
    end function
    integer(c_int) function simc_b200_fetch(h, acc) bind(C, name='simc_b200_fetch')
      import
      type(c_ptr), value :: h
      type(simc_accum), intent(inout) :: acc
    end function
    integer(c_int) function simc_b200_set_batch(h, tries_per_batch) bind(C, name='simc_b200_set_batch')
      import
      type(c_ptr), value :: h
      integer(c_int64_t), value :: tries_per_batch
    end function
    integer(c_int) function simc_b200_accum_merge(into, from) bind(C, name='simc_b200_accum_merge')
      import
      type(simc_accum), intent(inout) :: into
      type(simc_accum), intent(in) :: from
    end function
    ! ---- ntuple rows (results_ntu_write, results_write.f:1-269) and the .bin file (NtupleInit.f:32,352-355)
    integer(c_int) function simc_b200_ntuple_batch(h, first_try, n, seed, rows, n_cols, n_rows, try_of_row) &
        bind(C, name='simc_b200_ntuple_batch')
      import
      type(c_ptr), value :: h
      integer(c_int64_t), value :: first_try, n, seed
      real(c_double), intent(out) :: rows(*)
      integer(c_int32_t), intent(out) :: n_cols
      integer(c_int64_t), intent(out) :: n_rows
      type(c_ptr), value :: try_of_row
    end function
    integer(c_int) function simc_b200_ntuple_open(cfg, path, f) bind(C, name='simc_b200_ntuple_open')
      import
      type(simc_run_config), intent(in) :: cfg
      character(kind=c_char), intent(in) :: path(*)
      type(c_ptr), intent(out) :: f
    end function
    integer(c_int) function simc_b200_ntuple_append(f, rows, n_rows) bind(C, name='simc_b200_ntuple_append')
      import
      type(c_ptr), value :: f
      real(c_double), intent(in) :: rows(*)
      integer(c_int64_t), value :: n_rows
    end function
    integer(c_int) function simc_b200_ntuple_close(f) bind(C, name='simc_b200_ntuple_close')
      import
      type(c_ptr), value :: f
    end function
    ! ---- single-arm batch form of mc_hms / mc_shms / mc_sos / mc_hrsl / mc_hrsr (hms/mc_hms.f:1-4)
    integer(c_int) function simc_b200_transport_batch(h, arm, n, in_soa, seed, ms_flag, wcs_flag, decay_flag, using_coll, &
                                                      out_soa, flags) bind(C, name='simc_b200_transport_batch')
      import
      type(c_ptr), value :: h
      integer(c_int), value :: arm, ms_flag, wcs_flag, decay_flag, using_coll
      integer(c_int64_t), value :: n, seed
      real(c_double), intent(in) :: in_soa(*)
      real(c_double), intent(out) :: out_soa(*)
      integer(c_int32_t), intent(out) :: flags(*)
    end function
    ! ---- end of run (simc.f:94-101, 366-432)
    integer(c_int) function simc_b200_normalise(cfg, acc, ngen, charge_mC, res) bind(C, name='simc_b200_normalise')
      import
      type(simc_run_config), intent(in) :: cfg
      type(simc_accum), intent(in) :: acc
      integer(c_int32_t), value :: ngen
      real(c_double), value :: charge_mC
      type(simc_results), intent(out) :: res
    end function
  end interface

contains

  ! Fortran string -> NUL-terminated C string
  function c_path(s) result(c)
    character(len=*), intent(in) :: s
    character(kind=c_char) :: c(len_trim(s) + 1)
    integer :: i
    do i = 1, len_trim(s)
      c(i) = s(i:i)
    end do
    c(len_trim(s) + 1) = c_null_char
  end function

  ! value of a 128-bit fixed-point sum: (hi*2**64 + lo)*2**qexp with lo read as unsigned
  function simc_fixed_value(f) result(v)
    type(simc_fixed128), intent(in) :: f
    real(c_double) :: v
    integer, parameter :: xp = selected_real_kind(18)
    real(xp) :: lo_u, t
    lo_u = real(f%lo, xp)
    if (f%lo < 0) lo_u = lo_u + 18446744073709551616.0_xp
    t = real(f%hi, xp) * 18446744073709551616.0_xp + lo_u
    v = real(scale(t, int(f%qexp)), c_double)
  end function

end module simc_b200_api
