!  simc_b200_shim.f -- the event loop of `program simc` (simc.f:169-351) through libsimc_b200.
!
!  Fixed form like the rest of the reference (it includes the reference's own simulate.inc / radc.inc / struct_*.inc);
!  compile with the reference's flags (Makefile:63) next to simc_b200_api.f90 and link with -lsimc_b200.  In simc.f
!  the block from `nevent = 0` (simc.f:165) to `enddo ! <loop over ntried>` (simc.f:351) becomes
!
!        call simc_b200_loop(H, contrib, sumerr, sumerr2, sum_sigcc, central%sigcc)
!
!  and everything before (dbase_read, radc_init, calculate_central, luminosity) and after (normfac, resolutions,
!  .gen/.geni/.hist reports) stays as it is.  Three routines:
!    simc_b200_pack_run_config  COMMON /gnrl/ /radccom/ /target_info/ /decd/ + histogram axes -> simc_run_config
!    simc_b200_unpack_accum     simc_accum -> ntried, nevent, ncontribute, ..., H%*%buf, contrib, slop, *STOP_*
!    simc_b200_loop             create, load the tables of the reaction, run, unpack, destroy
!  This file cannot be compiled in the image the library was developed in (no Fortran compiler); the derived types
!  it relies on are checked against the C layout by tests/test_shim_layout.py.

	subroutine simc_b200_pack_run_config(cfg, H, w_ref)

	USE structureModule
	USE histoModule
	USE simc_b200_api
	implicit none
	include 'simulate.inc'
	include 'radc.inc'

	type(simc_run_config):: cfg
	type(histograms)::	H
	real*8			w_ref
	integer			i

	cfg%abi_version = SIMC_B200_ABI_VERSION
! ... reaction flags: /gnrl/ (simulate.inc:91-113), /decd/ (simulate.inc:153-158)
	cfg%doing_phsp      = merge(1,0,doing_phsp)
	cfg%doing_hyd_elast = merge(1,0,doing_hyd_elast)
	cfg%doing_deuterium = merge(1,0,doing_deuterium)
	cfg%doing_heavy     = merge(1,0,doing_heavy)
	cfg%doing_eep       = merge(1,0,doing_eep)
	cfg%doing_pion      = merge(1,0,doing_pion)
	cfg%doing_kaon      = merge(1,0,doing_kaon)
	cfg%doing_delta     = merge(1,0,doing_delta)
	cfg%doing_rho       = merge(1,0,doing_rho)
	cfg%doing_semi      = merge(1,0,doing_semi)
	cfg%doing_hydpi     = merge(1,0,doing_hydpi)
	cfg%doing_deutpi    = merge(1,0,doing_deutpi)
	cfg%doing_hepi      = merge(1,0,doing_hepi)
	cfg%doing_hydkaon   = merge(1,0,doing_hydkaon)
	cfg%doing_deutkaon  = merge(1,0,doing_deutkaon)
	cfg%doing_hekaon    = merge(1,0,doing_hekaon)
	cfg%doing_hydsemi   = merge(1,0,doing_hydsemi)
	cfg%doing_deutsemi  = merge(1,0,doing_deutsemi)
	cfg%doing_semipi    = merge(1,0,doing_semipi)
	cfg%doing_semika    = merge(1,0,doing_semika)
	cfg%do_fermi        = merge(1,0,do_fermi)
	cfg%doing_hplus     = merge(1,0,doing_hplus)
	cfg%doing_decay     = merge(1,0,doing_decay)
	cfg%which_pion      = which_pion
	cfg%which_kaon      = which_kaon
! ... switches, /gnrl/
	cfg%using_rad       = merge(1,0,using_rad)
	cfg%using_Eloss     = merge(1,0,using_Eloss)
	cfg%using_Coulomb   = merge(1,0,using_Coulomb)
	cfg%correct_Eloss   = merge(1,0,correct_Eloss)
	cfg%correct_raster  = merge(1,0,correct_raster)
	cfg%mc_smear        = merge(1,0,mc_smear)
	cfg%hard_cuts       = merge(1,0,hard_cuts)
	cfg%using_E_arm_montecarlo = merge(1,0,using_E_arm_montecarlo)
	cfg%using_P_arm_montecarlo = merge(1,0,using_P_arm_montecarlo)
	cfg%electron_arm    = electron_arm
	cfg%hadron_arm      = hadron_arm
	cfg%using_HMScoll   = merge(1,0,using_HMScoll)
	cfg%using_SHMScoll  = merge(1,0,using_SHMScoll)
	cfg%use_benhar_sf   = merge(1,0,use_benhar_sf)
! ... radiative flags, /radccom/ (radc.inc:13-19), as radc_init left them (init.f:621-629)
	cfg%rad_flag        = rad_flag
	cfg%extrad_flag     = extrad_flag
	cfg%intcor_mode     = intcor_mode
	cfg%use_expon       = use_expon
	cfg%use_offshell_rad = merge(1,0,use_offshell_rad)
	do i = 1, 3
	  cfg%doing_tail(i) = merge(1,0,doing_tail(i))
	enddo
	cfg%hardwired_rad   = merge(1,0,hardwired_rad)
	cfg%deForest_flag   = deForest_flag
	cfg%doing_pizero    = merge(1,0,doing_pizero)
	cfg%pizero_ngamma   = pizero_ngamma
	cfg%using_tgt_field = merge(1,0,using_tgt_field)
	cfg%pad_flags       = 0
! ... /gnrl/ scalars, ctau of /decd/
	cfg%Mh = Mh
	cfg%Mh2 = Mh2
	cfg%Ebeam = Ebeam
	cfg%dEbeam = dEbeam
	cfg%Ebeam_vertex_ave = Ebeam_vertex_ave
	cfg%dE_edge_test = dE_edge_test
	cfg%Egamma_gen_max = Egamma_gen_max
	cfg%ctau = ctau
	cfg%transparency = transparency
	cfg%drift_to_cal = drift_to_cal
	cfg%targ_Bangle = targ_Bangle
	cfg%targ_Bphi = targ_Bphi
	cfg%targ_pol = targ_pol
	cfg%sign_hadron = sign_hadron
! ... /radccom/ run-level
	cfg%etatzai = etatzai
	cfg%Egamma_tot_max = Egamma_tot_max
	cfg%Egamma1_max = Egamma1_max
	cfg%Egamma2_max = Egamma2_max
	cfg%Egamma3_max = Egamma3_max
	cfg%Egamma_res_limit = Egamma_res_limit
! ... gen (modules.f:210-215)
	cfg%gen%e%delta%min = gen%e%delta%min
	cfg%gen%e%delta%max = gen%e%delta%max
	cfg%gen%e%yptar%min = gen%e%yptar%min
	cfg%gen%e%yptar%max = gen%e%yptar%max
	cfg%gen%e%xptar%min = gen%e%xptar%min
	cfg%gen%e%xptar%max = gen%e%xptar%max
	cfg%gen%e%E%min = gen%e%E%min
	cfg%gen%e%E%max = gen%e%E%max
	cfg%gen%p%delta%min = gen%p%delta%min
	cfg%gen%p%delta%max = gen%p%delta%max
	cfg%gen%p%yptar%min = gen%p%yptar%min
	cfg%gen%p%yptar%max = gen%p%yptar%max
	cfg%gen%p%xptar%min = gen%p%xptar%min
	cfg%gen%p%xptar%max = gen%p%xptar%max
	cfg%gen%p%E%min = gen%p%E%min
	cfg%gen%p%E%max = gen%p%E%max
	cfg%gen%sumEgen%min = gen%sumEgen%min
	cfg%gen%sumEgen%max = gen%sumEgen%max
	cfg%gen%Trec%min = gen%Trec%min
	cfg%gen%Trec%max = gen%Trec%max
	cfg%gen%xwid = gen%xwid
	cfg%gen%ywid = gen%ywid
! ... spec (modules.f:150-165)
	cfg%spec_e%P = spec%e%P
	cfg%spec_e%theta = spec%e%theta
	cfg%spec_e%cos_th = spec%e%cos_th
	cfg%spec_e%sin_th = spec%e%sin_th
	cfg%spec_e%phi = spec%e%phi
	cfg%spec_e%off_x = spec%e%offset%x
	cfg%spec_e%off_y = spec%e%offset%y
	cfg%spec_e%off_z = spec%e%offset%z
	cfg%spec_e%off_xptar = spec%e%offset%xptar
	cfg%spec_e%off_yptar = spec%e%offset%yptar
	cfg%spec_p%P = spec%p%P
	cfg%spec_p%theta = spec%p%theta
	cfg%spec_p%cos_th = spec%p%cos_th
	cfg%spec_p%sin_th = spec%p%sin_th
	cfg%spec_p%phi = spec%p%phi
	cfg%spec_p%off_x = spec%p%offset%x
	cfg%spec_p%off_y = spec%p%offset%y
	cfg%spec_p%off_z = spec%p%offset%z
	cfg%spec_p%off_xptar = spec%p%offset%xptar
	cfg%spec_p%off_yptar = spec%p%offset%yptar
! ... cuts (modules.f:100-103)
	cfg%cuts_Em%min = cuts%Em%min
	cfg%cuts_Em%max = cuts%Em%max
	cfg%cuts_Pm%min = cuts%Pm%min
	cfg%cuts_Pm%max = cuts%Pm%max
! ... edge and VERTEXedge (modules.f:170-184)
	call simc_b200_pack_edge(cfg%edge, edge)
	call simc_b200_pack_edge(cfg%VERTEXedge, VERTEXedge)
! ... SPedge (modules.f:35-56)
	cfg%SPedge_e%delta%min = SPedge%e%delta%min
	cfg%SPedge_e%delta%max = SPedge%e%delta%max
	cfg%SPedge_e%yptar%min = SPedge%e%yptar%min
	cfg%SPedge_e%yptar%max = SPedge%e%yptar%max
	cfg%SPedge_e%xptar%min = SPedge%e%xptar%min
	cfg%SPedge_e%xptar%max = SPedge%e%xptar%max
	cfg%SPedge_e%z%min = SPedge%e%z%min
	cfg%SPedge_e%z%max = SPedge%e%z%max
	cfg%SPedge_p%delta%min = SPedge%p%delta%min
	cfg%SPedge_p%delta%max = SPedge%p%delta%max
	cfg%SPedge_p%yptar%min = SPedge%p%yptar%min
	cfg%SPedge_p%yptar%max = SPedge%p%yptar%max
	cfg%SPedge_p%xptar%min = SPedge%p%xptar%min
	cfg%SPedge_p%xptar%max = SPedge%p%xptar%max
	cfg%SPedge_p%z%min = SPedge%p%z%min
	cfg%SPedge_p%z%max = SPedge%p%z%max
! ... the slops pass_cuts reads (simc.f:219-240; modules.f:297-327)
	cfg%slop_MC_e_used(1) = slop%MC%e%delta%used
	cfg%slop_MC_e_used(2) = slop%MC%e%yptar%used
	cfg%slop_MC_e_used(3) = slop%MC%e%xptar%used
	cfg%slop_MC_p_used(1) = slop%MC%p%delta%used
	cfg%slop_MC_p_used(2) = slop%MC%p%yptar%used
	cfg%slop_MC_p_used(3) = slop%MC%p%xptar%used
! ... /target_info/ (target.inc:37-53)
	cfg%targ%A = targ%A
	cfg%targ%Z = targ%Z
	cfg%targ%N = targ%N
	cfg%targ%mass_amu = targ%mass_amu
	cfg%targ%M = targ%M
	cfg%targ%mrec_amu = targ%mrec_amu
	cfg%targ%Mrec = targ%Mrec
	cfg%targ%rho = targ%rho
	cfg%targ%thick = targ%thick
	cfg%targ%angle = targ%angle
	cfg%targ%abundancy = targ%abundancy
	cfg%targ%length = targ%length
	cfg%targ%zoffset = targ%zoffset
	cfg%targ%X0 = targ%X0
	cfg%targ%X0_cm = targ%X0_cm
	cfg%targ%L1 = targ%L1
	cfg%targ%L2 = targ%L2
	cfg%targ%fr1 = targ%fr1
	cfg%targ%fr2 = targ%fr2
	cfg%targ%xoffset = targ%xoffset
	cfg%targ%yoffset = targ%yoffset
	cfg%targ%Coulomb_ave = targ%Coulomb%ave
	cfg%targ%Coulomb_min = targ%Coulomb%min
	cfg%targ%Coulomb_max = targ%Coulomb%max
	cfg%targ%Coulomb_constant = targ%Coulomb_constant
	cfg%targ%Mtar_struck = targ%Mtar_struck
	cfg%targ%Mrec_struck = targ%Mrec_struck
	cfg%targ%fr_pattern = targ%fr_pattern
	cfg%targ%can = targ%can
! ... histogram axes (histograms_module.f:6-31, set up by init.f:519-569): hist_axis(slot,set), set 1 RECON, 2 gen, 3 geni
	call simc_b200_pack_axes(cfg%hist_axis(1,1), H%RECON)
	call simc_b200_pack_axes(cfg%hist_axis(1,2), H%gen)
	call simc_b200_pack_axes(cfg%hist_axis(1,3), H%geni)
! ... quantum of the exact weight sums: 2**(exponent(w_ref)-64); central%sigcc is the natural scale
	cfg%w_ref = w_ref

	return
	end

!-------------------------------------------------------------------

	subroutine simc_b200_pack_edge(c, e)

	USE structureModule
	USE simc_b200_api
	implicit none
	type(simc_edge)::	c
	type(edge_true)::	e

	c%e%E%min = e%e%E%min
	c%e%E%max = e%e%E%max
	c%e%yptar%min = e%e%yptar%min
	c%e%yptar%max = e%e%yptar%max
	c%e%xptar%min = e%e%xptar%min
	c%e%xptar%max = e%e%xptar%max
	c%p%E%min = e%p%E%min
	c%p%E%max = e%p%E%max
	c%p%yptar%min = e%p%yptar%min
	c%p%yptar%max = e%p%yptar%max
	c%p%xptar%min = e%p%xptar%min
	c%p%xptar%max = e%p%xptar%max
	c%Em%min = e%Em%min
	c%Em%max = e%Em%max
	c%Pm%min = e%Pm%min
	c%Pm%max = e%Pm%max
	c%Mrec%min = e%Mrec%min
	c%Mrec%max = e%Mrec%max
	c%Trec%min = e%Trec%min
	c%Trec%max = e%Trec%max
	c%Trec_struck%min = e%Trec_struck%min
	c%Trec_struck%max = e%Trec_struck%max
	return
	end

!-------------------------------------------------------------------

	subroutine simc_b200_pack_axes(ax, hs)

	USE histoModule
	USE simc_b200_api
	implicit none
	type(simc_axis)::	ax(8)
	type(hist_double_arm)::	hs

	ax(1)%min = hs%e%delta%min
	ax(1)%bin = hs%e%delta%bin
	ax(2)%min = hs%e%yptar%min
	ax(2)%bin = hs%e%yptar%bin
	ax(3)%min = hs%e%xptar%min
	ax(3)%bin = hs%e%xptar%bin
	ax(4)%min = hs%p%delta%min
	ax(4)%bin = hs%p%delta%bin
	ax(5)%min = hs%p%yptar%min
	ax(5)%bin = hs%p%yptar%bin
	ax(6)%min = hs%p%xptar%min
	ax(6)%bin = hs%p%xptar%bin
	ax(7)%min = hs%Em%min
	ax(7)%bin = hs%Em%bin
	ax(8)%min = hs%Pm%min
	ax(8)%bin = hs%Pm%bin
	return
	end

!-------------------------------------------------------------------
! What the loop body accumulates (simc.f:229-336) and what its callees count (the *STOP_* commons), from simc_accum.
! The histograms ADD (like `inc`), counters and sums ADD, ranges widen: the routine may be called once per chunk.

	subroutine simc_b200_unpack_accum(acc, H, contrib, sumerr, sumerr2, sum_sigcc)

	USE structureModule
	USE histoModule
	USE simc_b200_api
	implicit none
	include 'simulate.inc'
	include 'sos/struct_sos.inc'
	include 'hms/struct_hms.inc'
	include 'hrsr/struct_hrsr.inc'
	include 'hrsl/struct_hrsl.inc'
	include 'shms/struct_shms.inc'

	type(simc_accum)::	acc
	type(histograms)::	H
	type(contribtype)::	contrib
	type(sums_twoarm)::	sumerr, sumerr2
	real*8			sum_sigcc
	integer			i, w, arm

! ... counters (simulate.inc:60-61).  nevent: every try with ngen < 0, the successes with ngen > 0 (simc.f:346-350)
	ntried = ntried + int(acc%ntried)
	if (ngen.lt.0) then
	  nevent = nevent + int(acc%ntried)
	else
	  nevent = nevent + int(acc%nsuccess)
	endif
	ncontribute = ncontribute + int(acc%ncontribute)
	npasscuts = npasscuts + int(acc%npasscuts)
	ncontribute_no_rad_proton = ncontribute_no_rad_proton + int(acc%ncontribute_no_rad_proton)
	wtcontribute = wtcontribute + simc_fixed_value(acc%wtcontribute)
	sum_sigcc = sum_sigcc + simc_fixed_value(acc%sum_sigcc)
! ... reconstruction errors (simc.f:305-322): e delta, xptar, yptar, ytar; p the same
	sumerr%e%delta = sumerr%e%delta + simc_fixed_value(acc%sumerr(1))
	sumerr%e%xptar = sumerr%e%xptar + simc_fixed_value(acc%sumerr(2))
	sumerr%e%yptar = sumerr%e%yptar + simc_fixed_value(acc%sumerr(3))
	sumerr%e%ytar  = sumerr%e%ytar  + simc_fixed_value(acc%sumerr(4))
	sumerr%p%delta = sumerr%p%delta + simc_fixed_value(acc%sumerr(5))
	sumerr%p%xptar = sumerr%p%xptar + simc_fixed_value(acc%sumerr(6))
	sumerr%p%yptar = sumerr%p%yptar + simc_fixed_value(acc%sumerr(7))
	sumerr%p%ytar  = sumerr%p%ytar  + simc_fixed_value(acc%sumerr(8))
	sumerr2%e%delta = sumerr2%e%delta + simc_fixed_value(acc%sumerr2(1))
	sumerr2%e%xptar = sumerr2%e%xptar + simc_fixed_value(acc%sumerr2(2))
	sumerr2%e%yptar = sumerr2%e%yptar + simc_fixed_value(acc%sumerr2(3))
	sumerr2%e%ytar  = sumerr2%e%ytar  + simc_fixed_value(acc%sumerr2(4))
	sumerr2%p%delta = sumerr2%p%delta + simc_fixed_value(acc%sumerr2(5))
	sumerr2%p%xptar = sumerr2%p%xptar + simc_fixed_value(acc%sumerr2(6))
	sumerr2%p%yptar = sumerr2%p%yptar + simc_fixed_value(acc%sumerr2(7))
	sumerr2%p%ytar  = sumerr2%p%ytar  + simc_fixed_value(acc%sumerr2(8))
! ... histograms (simc.f:253-286): RECON spectrometer quantities are weighted, everything else counts
	do i = 1, nHbins
	  H%RECON%e%delta%buf(i) = H%RECON%e%delta%buf(i) + simc_fixed_value(acc%hist_w(i,1))
	  H%RECON%e%yptar%buf(i) = H%RECON%e%yptar%buf(i) + simc_fixed_value(acc%hist_w(i,2))
	  H%RECON%e%xptar%buf(i) = H%RECON%e%xptar%buf(i) + simc_fixed_value(acc%hist_w(i,3))
	  H%RECON%p%delta%buf(i) = H%RECON%p%delta%buf(i) + simc_fixed_value(acc%hist_w(i,4))
	  H%RECON%p%yptar%buf(i) = H%RECON%p%yptar%buf(i) + simc_fixed_value(acc%hist_w(i,5))
	  H%RECON%p%xptar%buf(i) = H%RECON%p%xptar%buf(i) + simc_fixed_value(acc%hist_w(i,6))
	  H%RECON%Em%buf(i) = H%RECON%Em%buf(i) + acc%hist_n(i,7,1)
	  H%RECON%Pm%buf(i) = H%RECON%Pm%buf(i) + acc%hist_n(i,8,1)
	  H%gen%e%delta%buf(i) = H%gen%e%delta%buf(i) + acc%hist_n(i,1,2)
	  H%gen%e%yptar%buf(i) = H%gen%e%yptar%buf(i) + acc%hist_n(i,2,2)
	  H%gen%e%xptar%buf(i) = H%gen%e%xptar%buf(i) + acc%hist_n(i,3,2)
	  H%gen%p%delta%buf(i) = H%gen%p%delta%buf(i) + acc%hist_n(i,4,2)
	  H%gen%p%yptar%buf(i) = H%gen%p%yptar%buf(i) + acc%hist_n(i,5,2)
	  H%gen%p%xptar%buf(i) = H%gen%p%xptar%buf(i) + acc%hist_n(i,6,2)
	  H%gen%Em%buf(i) = H%gen%Em%buf(i) + acc%hist_n(i,7,2)
	  H%geni%e%delta%buf(i) = H%geni%e%delta%buf(i) + acc%hist_n(i,1,3)
	  H%geni%e%yptar%buf(i) = H%geni%e%yptar%buf(i) + acc%hist_n(i,2,3)
	  H%geni%e%xptar%buf(i) = H%geni%e%xptar%buf(i) + acc%hist_n(i,3,3)
	  H%geni%p%delta%buf(i) = H%geni%p%delta%buf(i) + acc%hist_n(i,4,3)
	  H%geni%p%yptar%buf(i) = H%geni%p%yptar%buf(i) + acc%hist_n(i,5,3)
	  H%geni%p%xptar%buf(i) = H%geni%p%xptar%buf(i) + acc%hist_n(i,6,3)
	  H%geni%Em%buf(i) = H%geni%Em%buf(i) + acc%hist_n(i,7,3)
	  H%geni%Pm%buf(i) = H%geni%Pm%buf(i) + acc%hist_n(i,8,3)
	enddo
! ... contribution ranges, in the order of limits_update (event.f:19-72; the second contrib%tru%e%xptar call of
! ... event.f:41 repeats event.f:39 and has no slot of its own)
	call simc_b200_widen(contrib%gen%e%delta, acc%contrib(1))
	call simc_b200_widen(contrib%gen%e%yptar, acc%contrib(2))
	call simc_b200_widen(contrib%gen%e%xptar, acc%contrib(3))
	call simc_b200_widen(contrib%gen%p%delta, acc%contrib(4))
	call simc_b200_widen(contrib%gen%p%yptar, acc%contrib(5))
	call simc_b200_widen(contrib%gen%p%xptar, acc%contrib(6))
	call simc_b200_widen(contrib%gen%Trec, acc%contrib(7))
	call simc_b200_widen(contrib%gen%sumEgen, acc%contrib(8))
	call simc_b200_widen(contrib%tru%e%E, acc%contrib(9))
	call simc_b200_widen(contrib%tru%e%xptar, acc%contrib(10))
	call simc_b200_widen(contrib%tru%e%yptar, acc%contrib(11))
	call simc_b200_widen(contrib%tru%p%E, acc%contrib(12))
	call simc_b200_widen(contrib%tru%p%yptar, acc%contrib(13))
	call simc_b200_widen(contrib%tru%p%xptar, acc%contrib(14))
	call simc_b200_widen(contrib%tru%Em, acc%contrib(15))
	call simc_b200_widen(contrib%tru%Pm, acc%contrib(16))
	call simc_b200_widen(contrib%tru%Trec, acc%contrib(17))
	call simc_b200_widen(contrib%SP%e%delta, acc%contrib(18))
	call simc_b200_widen(contrib%SP%e%yptar, acc%contrib(19))
	call simc_b200_widen(contrib%SP%e%xptar, acc%contrib(20))
	call simc_b200_widen(contrib%SP%p%delta, acc%contrib(21))
	call simc_b200_widen(contrib%SP%p%yptar, acc%contrib(22))
	call simc_b200_widen(contrib%SP%p%xptar, acc%contrib(23))
	call simc_b200_widen(contrib%vertex%Trec, acc%contrib(24))
	call simc_b200_widen(contrib%vertex%Em, acc%contrib(25))
	call simc_b200_widen(contrib%vertex%Pm, acc%contrib(26))
	call simc_b200_widen(contrib%rad%Egamma(1), acc%contrib(27))
	call simc_b200_widen(contrib%rad%Egamma(2), acc%contrib(28))
	call simc_b200_widen(contrib%rad%Egamma(3), acc%contrib(29))
	call simc_b200_widen(contrib%rad%Egamma_total, acc%contrib(30))
! ... slop ranges (event.f:75-87)
	slop%MC%e%delta%lo = min(slop%MC%e%delta%lo, acc%slop(1)%lo)
	slop%MC%e%delta%hi = max(slop%MC%e%delta%hi, acc%slop(1)%hi)
	slop%MC%e%yptar%lo = min(slop%MC%e%yptar%lo, acc%slop(2)%lo)
	slop%MC%e%yptar%hi = max(slop%MC%e%yptar%hi, acc%slop(2)%hi)
	slop%MC%e%xptar%lo = min(slop%MC%e%xptar%lo, acc%slop(3)%lo)
	slop%MC%e%xptar%hi = max(slop%MC%e%xptar%hi, acc%slop(3)%hi)
	slop%MC%p%delta%lo = min(slop%MC%p%delta%lo, acc%slop(4)%lo)
	slop%MC%p%delta%hi = max(slop%MC%p%delta%hi, acc%slop(4)%hi)
	slop%MC%p%yptar%lo = min(slop%MC%p%yptar%lo, acc%slop(5)%lo)
	slop%MC%p%yptar%hi = max(slop%MC%p%yptar%hi, acc%slop(5)%hi)
	slop%MC%p%xptar%lo = min(slop%MC%p%xptar%lo, acc%slop(6)%lo)
	slop%MC%p%xptar%hi = max(slop%MC%p%xptar%hi, acc%slop(6)%hi)
	slop%total%Em%lo = min(slop%total%Em%lo, acc%slop(7)%lo)
	slop%total%Em%hi = max(slop%total%Em%hi, acc%slop(7)%hi)
	slop%total%Pm%lo = min(slop%total%Pm%lo, acc%slop(8)%lo)
	slop%total%Pm%hi = max(slop%total%Pm%hi, acc%slop(8)%hi)
! ... where events were lost: stop(slot,w), w = 1 electron arm, 2 hadron arm; slot 1 trials, 2 successes, 3 events
! ... reaching the hut, 3 + code for the apertures in the order of simc_b200_stop_name (code 1 = first aperture)
	do w = 1, 2
	  arm = electron_arm
	  if (w.eq.2) arm = hadron_arm
	  if (arm.eq.1) then
	    hSTOP_trials    = hSTOP_trials    + int(acc%stop(1,w))
	    hSTOP_successes = hSTOP_successes + int(acc%stop(2,w))
	    hSTOP_hut       = hSTOP_hut       + int(acc%stop(3,w))
	    hSTOP_slit_hor  = hSTOP_slit_hor  + int(acc%stop(4,w))
	    hSTOP_slit_vert = hSTOP_slit_vert + int(acc%stop(5,w))
	    hSTOP_slit_oct  = hSTOP_slit_oct  + int(acc%stop(6,w))
	    hSTOP_Q1_in     = hSTOP_Q1_in     + int(acc%stop(7,w))
	    hSTOP_Q1_mid    = hSTOP_Q1_mid    + int(acc%stop(8,w))
	    hSTOP_Q1_out    = hSTOP_Q1_out    + int(acc%stop(9,w))
	    hSTOP_Q2_in     = hSTOP_Q2_in     + int(acc%stop(10,w))
	    hSTOP_Q2_mid    = hSTOP_Q2_mid    + int(acc%stop(11,w))
	    hSTOP_Q2_out    = hSTOP_Q2_out    + int(acc%stop(12,w))
	    hSTOP_Q3_in     = hSTOP_Q3_in     + int(acc%stop(13,w))
	    hSTOP_Q3_mid    = hSTOP_Q3_mid    + int(acc%stop(14,w))
	    hSTOP_Q3_out    = hSTOP_Q3_out    + int(acc%stop(15,w))
	    hSTOP_D1_in     = hSTOP_D1_in     + int(acc%stop(16,w))
	    hSTOP_D1_out    = hSTOP_D1_out    + int(acc%stop(17,w))
	    hSTOP_dc1       = hSTOP_dc1       + int(acc%stop(18,w))
	    hSTOP_dc2       = hSTOP_dc2       + int(acc%stop(19,w))
	    hSTOP_scin      = hSTOP_scin      + int(acc%stop(20,w))
	    hSTOP_cal       = hSTOP_cal       + int(acc%stop(21,w))
	    hSTOP_coll      = hSTOP_coll      + int(acc%stop(22,w))
	  else if (arm.eq.2) then
	    sSTOP_trials    = sSTOP_trials    + int(acc%stop(1,w))
	    sSTOP_successes = sSTOP_successes + int(acc%stop(2,w))
	    sSTOP_hut       = sSTOP_hut       + int(acc%stop(3,w))
	    sSTOP_slit_hor  = sSTOP_slit_hor  + int(acc%stop(4,w))
	    sSTOP_slit_vert = sSTOP_slit_vert + int(acc%stop(5,w))
	    sSTOP_slit_oct  = sSTOP_slit_oct  + int(acc%stop(6,w))
	    sSTOP_quad_in   = sSTOP_quad_in   + int(acc%stop(7,w))
	    sSTOP_quad_mid  = sSTOP_quad_mid  + int(acc%stop(8,w))
	    sSTOP_quad_out  = sSTOP_quad_out  + int(acc%stop(9,w))
	    sSTOP_bm01_in   = sSTOP_bm01_in   + int(acc%stop(10,w))
	    sSTOP_bm01_out  = sSTOP_bm01_out  + int(acc%stop(11,w))
	    sSTOP_bm02_in   = sSTOP_bm02_in   + int(acc%stop(12,w))
	    sSTOP_bm02_out  = sSTOP_bm02_out  + int(acc%stop(13,w))
	    sSTOP_exit      = sSTOP_exit      + int(acc%stop(14,w))
	    sSTOP_dc1       = sSTOP_dc1       + int(acc%stop(15,w))
	    sSTOP_dc2       = sSTOP_dc2       + int(acc%stop(16,w))
	    sSTOP_scin      = sSTOP_scin      + int(acc%stop(17,w))
	  else if (arm.eq.3) then
	    rSTOP_trials    = rSTOP_trials    + int(acc%stop(1,w))
	    rSTOP_successes = rSTOP_successes + int(acc%stop(2,w))
	    rSTOP_hut       = rSTOP_hut       + int(acc%stop(3,w))
	    rSTOP_slit_hor  = rSTOP_slit_hor  + int(acc%stop(4,w))
	    rSTOP_slit_vert = rSTOP_slit_vert + int(acc%stop(5,w))
	    rSTOP_Q1_in     = rSTOP_Q1_in     + int(acc%stop(6,w))
	    rSTOP_Q1_mid    = rSTOP_Q1_mid    + int(acc%stop(7,w))
	    rSTOP_Q1_out    = rSTOP_Q1_out    + int(acc%stop(8,w))
	    rSTOP_Q2_in     = rSTOP_Q2_in     + int(acc%stop(9,w))
	    rSTOP_Q2_mid    = rSTOP_Q2_mid    + int(acc%stop(10,w))
	    rSTOP_Q2_out    = rSTOP_Q2_out    + int(acc%stop(11,w))
	    rSTOP_D1_in     = rSTOP_D1_in     + int(acc%stop(12,w))
	    rSTOP_D1_out    = rSTOP_D1_out    + int(acc%stop(13,w))
	    rSTOP_Q3_in     = rSTOP_Q3_in     + int(acc%stop(14,w))
	    rSTOP_Q3_mid    = rSTOP_Q3_mid    + int(acc%stop(15,w))
	    rSTOP_Q3_out    = rSTOP_Q3_out    + int(acc%stop(16,w))
	    rSTOP_dc1       = rSTOP_dc1       + int(acc%stop(17,w))
	    rSTOP_dc2       = rSTOP_dc2       + int(acc%stop(18,w))
	    rSTOP_s1        = rSTOP_s1        + int(acc%stop(19,w))
	    rSTOP_s2        = rSTOP_s2        + int(acc%stop(20,w))
	  else if (arm.eq.4) then
	    lSTOP_trials    = lSTOP_trials    + int(acc%stop(1,w))
	    lSTOP_successes = lSTOP_successes + int(acc%stop(2,w))
	    lSTOP_hut       = lSTOP_hut       + int(acc%stop(3,w))
	    lSTOP_slit_hor  = lSTOP_slit_hor  + int(acc%stop(4,w))
	    lSTOP_slit_vert = lSTOP_slit_vert + int(acc%stop(5,w))
	    lSTOP_Q1_in     = lSTOP_Q1_in     + int(acc%stop(6,w))
	    lSTOP_Q1_mid    = lSTOP_Q1_mid    + int(acc%stop(7,w))
	    lSTOP_Q1_out    = lSTOP_Q1_out    + int(acc%stop(8,w))
	    lSTOP_Q2_in     = lSTOP_Q2_in     + int(acc%stop(9,w))
	    lSTOP_Q2_mid    = lSTOP_Q2_mid    + int(acc%stop(10,w))
	    lSTOP_Q2_out    = lSTOP_Q2_out    + int(acc%stop(11,w))
	    lSTOP_D1_in     = lSTOP_D1_in     + int(acc%stop(12,w))
	    lSTOP_D1_out    = lSTOP_D1_out    + int(acc%stop(13,w))
	    lSTOP_Q3_in     = lSTOP_Q3_in     + int(acc%stop(14,w))
	    lSTOP_Q3_mid    = lSTOP_Q3_mid    + int(acc%stop(15,w))
	    lSTOP_Q3_out    = lSTOP_Q3_out    + int(acc%stop(16,w))
	    lSTOP_dc1       = lSTOP_dc1       + int(acc%stop(17,w))
	    lSTOP_dc2       = lSTOP_dc2       + int(acc%stop(18,w))
	    lSTOP_s1        = lSTOP_s1        + int(acc%stop(19,w))
	    lSTOP_s2        = lSTOP_s2        + int(acc%stop(20,w))
	  else if (arm.eq.5 .or. arm.eq.6) then
	    shmsSTOP_trials    = shmsSTOP_trials    + int(acc%stop(1,w))
	    shmsSTOP_successes = shmsSTOP_successes + int(acc%stop(2,w))
	    shmsSTOP_hut       = shmsSTOP_hut       + int(acc%stop(3,w))
	    shmsSTOP_HB_in     = shmsSTOP_HB_in     + int(acc%stop(4,w))
	    shmsSTOP_HB_men    = shmsSTOP_HB_men    + int(acc%stop(5,w))
	    shmsSTOP_HB_mex    = shmsSTOP_HB_mex    + int(acc%stop(6,w))
	    shmsSTOP_HB_out    = shmsSTOP_HB_out    + int(acc%stop(7,w))
	    shmsSTOP_slit_hor  = shmsSTOP_slit_hor  + int(acc%stop(8,w))
	    shmsSTOP_slit_vert = shmsSTOP_slit_vert + int(acc%stop(9,w))
	    shmsSTOP_slit_oct  = shmsSTOP_slit_oct  + int(acc%stop(10,w))
	    shmsSTOP_Q1_in     = shmsSTOP_Q1_in     + int(acc%stop(11,w))
	    shmsSTOP_Q1_men    = shmsSTOP_Q1_men    + int(acc%stop(12,w))
	    shmsSTOP_Q1_mid    = shmsSTOP_Q1_mid    + int(acc%stop(13,w))
	    shmsSTOP_Q1_mex    = shmsSTOP_Q1_mex    + int(acc%stop(14,w))
	    shmsSTOP_Q1_out    = shmsSTOP_Q1_out    + int(acc%stop(15,w))
	    shmsSTOP_Q2_in     = shmsSTOP_Q2_in     + int(acc%stop(16,w))
	    shmsSTOP_Q2_men    = shmsSTOP_Q2_men    + int(acc%stop(17,w))
	    shmsSTOP_Q2_mid    = shmsSTOP_Q2_mid    + int(acc%stop(18,w))
	    shmsSTOP_Q2_mex    = shmsSTOP_Q2_mex    + int(acc%stop(19,w))
	    shmsSTOP_Q2_out    = shmsSTOP_Q2_out    + int(acc%stop(20,w))
	    shmsSTOP_Q3_in     = shmsSTOP_Q3_in     + int(acc%stop(21,w))
	    shmsSTOP_Q3_men    = shmsSTOP_Q3_men    + int(acc%stop(22,w))
	    shmsSTOP_Q3_mid    = shmsSTOP_Q3_mid    + int(acc%stop(23,w))
	    shmsSTOP_Q3_mex    = shmsSTOP_Q3_mex    + int(acc%stop(24,w))
	    shmsSTOP_Q3_out    = shmsSTOP_Q3_out    + int(acc%stop(25,w))
	    shmsSTOP_D1_in     = shmsSTOP_D1_in     + int(acc%stop(26,w))
	    shmsSTOP_D1_flr    = shmsSTOP_D1_flr    + int(acc%stop(27,w))
	    shmsSTOP_D1_men    = shmsSTOP_D1_men    + int(acc%stop(28,w))
	    shmsSTOP_D1_mid1   = shmsSTOP_D1_mid1   + int(acc%stop(29,w))
	    shmsSTOP_D1_mid2   = shmsSTOP_D1_mid2   + int(acc%stop(30,w))
	    shmsSTOP_D1_mid3   = shmsSTOP_D1_mid3   + int(acc%stop(31,w))
	    shmsSTOP_D1_mid4   = shmsSTOP_D1_mid4   + int(acc%stop(32,w))
	    shmsSTOP_D1_mid5   = shmsSTOP_D1_mid5   + int(acc%stop(33,w))
	    shmsSTOP_D1_mid6   = shmsSTOP_D1_mid6   + int(acc%stop(34,w))
	    shmsSTOP_D1_mid7   = shmsSTOP_D1_mid7   + int(acc%stop(35,w))
	    shmsSTOP_D1_mex    = shmsSTOP_D1_mex    + int(acc%stop(36,w))
	    shmsSTOP_D1_out    = shmsSTOP_D1_out    + int(acc%stop(37,w))
	    shmsSTOP_dc1       = shmsSTOP_dc1       + int(acc%stop(38,w))
	    shmsSTOP_dc2       = shmsSTOP_dc2       + int(acc%stop(39,w))
! ... the hut counts both S1 planes in shmsSTOP_s1, S2X in shmsSTOP_s3 and S2Y in shmsSTOP_s2, and both
! ... calorimeter tests in shmsSTOP_cal (mc_shms_hut.f:298,313,359,374,396,439)
	    shmsSTOP_s1        = shmsSTOP_s1        + int(acc%stop(40,w)) + int(acc%stop(41,w))
	    shmsSTOP_s3        = shmsSTOP_s3        + int(acc%stop(42,w))
	    shmsSTOP_s2        = shmsSTOP_s2        + int(acc%stop(43,w))
	    shmsSTOP_cal       = shmsSTOP_cal       + int(acc%stop(44,w)) + int(acc%stop(45,w))
	    sSTOP_coll         = sSTOP_coll         + int(acc%stop(46,w))
	  endif
	enddo

	return
	end

!-------------------------------------------------------------------

	subroutine simc_b200_widen(range, r)

	USE structureModule
	USE simc_b200_api
	implicit none
	type(rangetype)::	range
	type(simc_range)::	r

	range%lo = min(range%lo, r%lo)
	range%hi = max(range%hi, r%hi)
	return
	end

!-------------------------------------------------------------------
! The loop.  ngen < 0: |ngen| tries.  ngen > 0: until ngen successes; the try range of the last chunk is bisected so
! that the run ends with the try that gave the ngen-th success (simc.f:346-350; every try is reproducible from
! (random_seed, try index)).  Ntuple rows (Nntu > 0) are written through results_ntu_write's unit by the library's
! row producer.

	subroutine simc_b200_loop(H, contrib, sumerr, sumerr2, sum_sigcc, w_ref)

	USE structureModule
	USE histoModule
	USE simc_b200_api
	implicit none
	include 'simulate.inc'
	include 'radc.inc'
	include 'hbook.inc'

	type(histograms)::	H
	type(contribtype)::	contrib
	type(sums_twoarm)::	sumerr, sumerr2
	real*8			sum_sigcc, w_ref

	type(simc_run_config)::	cfg
	type(simc_accum), target:: acc, trial
	type(c_ptr)::		hnd
	integer(c_int)::	ierr
	integer(c_int64_t)::	first, n, chunk, lo, hi, mid, seed, need, n_rows
	integer(c_int32_t)::	n_cols
	real(c_double), allocatable:: rows(:)
	integer			i, k
	character*80		fwd, rec

	chunk = 4194304
	seed = random_seed
	call simc_b200_pack_run_config(cfg, H, w_ref)
	ierr = simc_b200_create(cfg, 0_c_int, hnd)
	if (ierr.ne.0) stop 'simc_b200_create failed'
! ... COSY tables: the files transp_init and mc_*_recon open (shared/transp.f:294-474, hms/mc_hms_recon.f:70-102)
	do k = 1, 2
	  i = electron_arm
	  if (k.eq.2) i = hadron_arm
	  if (i.eq.1) then
	    fwd = 'hms/forward_cosy.dat'
	    rec = 'hms/recon_cosy.dat'
	  else if (i.eq.2) then
	    fwd = 'sos/forward_cosy.dat'
	    rec = 'sos/recon_cosy.dat'
	  else if (i.eq.3) then
	    fwd = 'hrsr/hrs_forward_cosy.dat'
	    rec = 'hrsr/hrs_recon_cosy.dat'
	  else if (i.eq.4) then
	    fwd = 'hrsl/hrs_forward_cosy.dat'
	    rec = 'hrsl/hrs_recon_cosy.dat'
	  else
	    fwd = 'shms/shms_forward.dat'
	    rec = 'shms/shms_recon.dat'
	  endif
	  ierr = simc_b200_load_optics(hnd, int(i,c_int), c_path(fwd), c_path(rec))
	  if (ierr.ne.0) stop 'simc_b200_load_optics failed'
	enddo
! ... tables of the reaction (INTEGRATION.md has the list); the loaders read the reference's own files
	if (doing_heavy .and. use_benhar_sf) ierr = simc_b200_load_sf_file(hnd, c_path('benharsf_12.dat'), 1_c_int)
	if (doing_deuterium .or. (doing_heavy .and. .not.use_benhar_sf))
     >		ierr = simc_b200_load_theory_file(hnd, c_path(theory_file))
	if (doing_semi) ierr = simc_b200_load_cteq5_file(hnd, c_path('cteq5/cteq5m.tbl'))
	if (doing_semi .and. doing_semika) ierr = simc_b200_load_fdss_file(hnd, c_path('fdss/KANLO.GRID'))
	if (doing_deutsemi .or. doing_deutpi .or. doing_deutkaon)
     >		ierr = simc_b200_load_pfermi_file(hnd, c_path('deut.dat'))
	if (doing_pion .and. (which_pion.eq.0 .or. which_pion.eq.10 .or. which_pion.eq.2))
     >		ierr = simc_b200_load_maid_file(hnd, 3_c_int, c_path('maidpipn.dat'))
	if (doing_pion .and. (which_pion.eq.1 .or. which_pion.eq.11 .or. which_pion.eq.3))
     >		ierr = simc_b200_load_maid_file(hnd, 4_c_int, c_path('maidpimp.dat'))
	if (doing_kaon) ierr = simc_b200_load_saghai_files(hnd, c_path('.'))
! ... the field map trgInit reads (simc.f:154); the angles between the field axis and the two arms (simc.f:120-152) are
! ... rebuilt by the library from targ_Bangle, targ_Bphi and the spectrometer angles it was given
	if (using_tgt_field) ierr = simc_b200_load_field_file(hnd, c_path(trim(tgt_field_file)))
	if (Nntu.gt.0) allocate(rows(chunk*SIMC_NTUPLE_MAXCOL))

	first = 0
	nevent = 0
	do while (nevent.lt.abs(ngen))
	  ierr = simc_b200_accum_clear(hnd, acc)
	  if (ngen.lt.0) then
	    n = min(chunk, int(abs(ngen),c_int64_t) - first)
	    ierr = simc_b200_run(hnd, first, n, seed, acc)
	    if (ierr.ne.0) stop 'simc_b200_run failed'
	  else
	    n = chunk
	    need = ngen - nevent
	    ierr = simc_b200_run(hnd, first, n, seed, acc)
	    if (ierr.ne.0) stop 'simc_b200_run failed'
	    if (acc%nsuccess .ge. need) then
	      lo = 0
	      hi = n
	      do while (hi-lo .gt. 1)
	        mid = (lo+hi)/2
	        ierr = simc_b200_accum_clear(hnd, trial)
	        ierr = simc_b200_run(hnd, first, mid, seed, trial)
	        if (trial%nsuccess .ge. need) then
	          hi = mid
	        else
	          lo = mid
	        endif
	      enddo
	      n = hi
	      ierr = simc_b200_accum_clear(hnd, acc)
	      ierr = simc_b200_run(hnd, first, n, seed, acc)
	    endif
	  endif
	  if (Nntu.gt.0) then
	    ierr = simc_b200_ntuple_batch(hnd, first, n, seed, rows, n_cols, n_rows, c_null_ptr)
	    do i = 1, int(n_rows)
	      do k = 1, n_cols
	        write(NtupleIO) rows((i-1)*SIMC_NTUPLE_MAXCOL + k)
	      enddo
	    enddo
	  endif
	  call simc_b200_unpack_accum(acc, H, contrib, sumerr, sumerr2, sum_sigcc)
	  first = first + n
	enddo
	call simc_b200_destroy(hnd)

	return
	end
