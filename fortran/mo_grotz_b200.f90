!>
!! Drop-in replacement of mo_grotz.f90: same module name, same entry `grotz(testcase, description)`, same mo_data /
!! mo_parameters state, same output files -- the time loop (mo_grotz.f90:182-835) runs on the B200 through
!! mo_samsim_b200.  `grotz_batch` is the batched entry (ensembles, perturbed forcing); grotz == grotz_batch with
!! one column.
!!
!! What stays from the reference, unchanged: init (mo_init.f90:73), output_begin/output_settings/output
!! (mo_output.f90), sub_input (mo_functions.f90:304), the lab READs (mo_grotz.f90:138-169), the closing block
!! (mo_grotz.f90:840-876).  What goes: the `DO i = 1,i_time` body.
!!
!! NOTE: shipped uncompiled (no Fortran compiler in this image); see INTEGRATION.md.
!!
MODULE mo_grotz

  USE, INTRINSIC :: ISO_C_BINDING
  IMPLICIT NONE
  PUBLIC :: grotz, grotz_batch

CONTAINS

  SUBROUTINE grotz(testcase, description)
    INTEGER,         INTENT(in) :: testcase
    CHARACTER*12000, INTENT(in) :: description
    CALL grotz_batch(testcase, description, 1)
  END SUBROUTINE grotz

  !> ncol identical columns are advanced (perturb them between b200_push_mo_data and the loop through
  !! samsim_b200_set_scalar / samsim_b200_set_forcing's scale and offset vectors); column 1 drives the output files.
  SUBROUTINE grotz_batch(testcase, description, ncol)
    USE mo_parameters
    USE mo_data
    USE mo_init
    USE mo_output
    USE mo_functions
    USE mo_samsim_b200

    INTEGER,         INTENT(in) :: testcase, ncol
    CHARACTER*12000, INTENT(in) :: description

    TYPE(C_PTR)            :: h
    TYPE(samsim_config_t)  :: cfg
    INTEGER(C_INT64_T)     :: n, done, total
    INTEGER(C_INT32_T)     :: col_status(1)
    LOGICAL                :: wrote                 ! (the layer index k is mo_data's)
    REAL(C_DOUBLE), ALLOCATABLE :: series(:), snap_sc(:), snap_ar(:,:)
    REAL(wp), ALLOCATABLE  :: o_T(:), o_psi_s(:), o_thick(:), o_S_bu(:), o_ray(:), o_psi_l(:), o_perm(:), o_fv(:), &
         &                    o_fh(:), o_psi_g(:)
    REAL(wp)               :: o_mto(3)
    CHARACTER(len=8)       :: fmt = '(I1)'
    CHARACTER(len=1)       :: num

    !---- unchanged reference prologue (mo_grotz.f90:119-176) -------------------------------------------------
    CALL init(testcase)
    CALL output_begin(Nlayer,debug_flag,format_T,format_psi,format_thick,format_snow,format_T2m_top,format_perm,format_melt)
    CALL output_settings(description,testcase,N_top,N_bottom,Nlayer,fl_q_bottom,T_bottom,S_bu_bottom,thick_0,time_out, &
         & time_total,dt,boundflux_flag,atmoflux_flag,albedo_flag,grav_flag,flush_flag,flood_flag,grav_heat_flag,      &
         & flush_heat_flag,harmonic_flag,prescribe_flag,salt_flag,turb_flag,bottom_flag,tank_flag,precip_flag,bgc_flag, &
         & N_bgc,k_snow_flush)
    IF (bgc_flag == 2) CALL output_begin_bgc(Nlayer,N_bgc,format_bgc)      ! mo_grotz.f90:121-123
    IF (atmoflux_flag == 2) THEN
       Length_Input = 13148
       time_counter = 1
       CALL sub_input(length_input,fl_sw_input,fl_lw_input,T2m_input,precip_input,time_input)
    END IF
    IF (testcase >= 101 .AND. testcase <= 105) THEN
       WRITE(num,fmt) testcase-100
       OPEN(1234,file='2017_input/Tice_exp_'//num//'.txt',status='old');     READ(1234,*) Tinput;           CLOSE(1234)
       OPEN(1235,file='2017_input/snowfall_exp_'//num//'.txt',status='old'); READ(1235,*) precipinput;      CLOSE(1235)
       OPEN(1235,file='2017_input/heat_exp_'//num//'.txt',status='old');     READ(1235,*) ocean_flux_input; CLOSE(1235)
       OPEN(1234,file='2017_input/styropor_exp_'//num//'.txt',status='old'); READ(1234,*) styropor_input;   CLOSE(1234)
    END IF

    IF (testcase == 111) THEN                     ! mo_grotz.f90:171-176
       WRITE(num,fmt) INT(dt)
       OPEN(1234,file='2017_input/Ts_'//num//'s.txt',status='old'); READ(1234,*) Ttop_input; CLOSE(1234)
    END IF

    !---- device set-up -----------------------------------------------------------------------------------------
    i = 1                                   ! mo_data loop index: the next step to execute
    cfg = b200_config_from_mo_data(testcase)
    CALL b200_check(samsim_b200_create(cfg, INT(ncol, C_INT32_T), 0_C_INT32_T, h), 'create')
    CALL b200_push_mo_data(h)
    IF (ncol > 1) CALL b200_check(samsim_b200_broadcast_column(h, 0_C_INT32_T, 0_C_INT32_T, INT(ncol, C_INT32_T)), 'broadcast')
    IF (atmoflux_flag == 2) THEN            ! series[(site*4+kind)*nrec + r], kinds fl_sw, fl_lw, T2m, precip
       ALLOCATE(series(4*Length_Input))
       series(1:Length_Input)                  = fl_sw_input
       series(Length_Input+1:2*Length_Input)   = fl_lw_input
       series(2*Length_Input+1:3*Length_Input) = T2m_input
       series(3*Length_Input+1:4*Length_Input) = precip_input
       CALL b200_check(samsim_b200_set_forcing(h, 1_C_INT32_T, INT(Length_Input, C_INT32_T), series, C_NULL_PTR, C_NULL_PTR, &
            &                                  C_NULL_PTR), 'set_forcing')
       DEALLOCATE(series)
    END IF
    IF (testcase >= 101 .AND. testcase <= 105) THEN   ! kinds Tice, snowfall, heat, styropor
       ALLOCATE(series(4*length_input_lab))
       series(1:length_input_lab)                      = Tinput
       series(length_input_lab+1:2*length_input_lab)   = precipinput
       series(2*length_input_lab+1:3*length_input_lab) = ocean_flux_input
       series(3*length_input_lab+1:4*length_input_lab) = styropor_input
       CALL b200_check(samsim_b200_set_lab_forcing(h, 1_C_INT32_T, INT(length_input_lab, C_INT64_T), series, C_NULL_PTR), 'lab')
       DEALLOCATE(series)
    END IF
    IF (testcase == 111) THEN                         ! T_top = Ttop_input(FLOOR(1+time/dt)): the series travels as kind Tice
       ALLOCATE(series(4*length_input_lab))
       series = 0.0_C_DOUBLE
       series(1:length_input_lab) = Ttop_input
       CALL b200_check(samsim_b200_set_lab_forcing(h, 1_C_INT32_T, INT(length_input_lab, C_INT64_T), series, C_NULL_PTR), 'lab')
       DEALLOCATE(series)
    END IF
    CALL b200_check(samsim_b200_set_snapshot_mode(h, SAMSIM_SNAP_FULL), 'snapshot mode')
    ALLOCATE(snap_sc(20), snap_ar(Nlayer,14))   ! SAMSIM_SNAPSC_COUNT, SAMSIM_SNAPARR_COUNT of include/samsim_b200.h
    ALLOCATE(o_T(Nlayer), o_psi_s(Nlayer), o_thick(Nlayer), o_S_bu(Nlayer), o_ray(Nlayer-1), o_psi_l(Nlayer), &
         &   o_perm(Nlayer), o_fv(Nlayer), o_fh(Nlayer), o_psi_g(Nlayer))

    !---- the time loop: DO i = 1,i_time (mo_grotz.f90:182) in chunks that end on output steps -------------------
    total = INT(i_time, C_INT64_T)
    done  = 0
    DO WHILE (done < total)
       n     = samsim_b200_steps_to_next_output(h)
       IF (n <= 0) THEN                      ! cannot happen with a consistent clock; never spin on step(h, 0)
          WRITE(*,*) 'samsim_b200: steps_to_next_output returned ', n
          STOP 1
       END IF
       wrote = (done + n <= total)
       n     = MIN(n, total - done)
       CALL b200_check(samsim_b200_step(h, n), 'step')
       done = done + n
       ! the reference STOPs the moment a check fails (codes 99, 16, 345, 9876, 21234, 1337, 431, 7889); column 1 is
       ! the drop-in column, so its status ends the run here instead of after the loop
       CALL b200_check(samsim_b200_get_status(h, col_status, 0_C_INT32_T, 1_C_INT32_T), 'status')
       IF (col_status(1) /= 0) THEN
          WRITE(*,*) 'samsim_b200: column 1 stopped with the reference code ', col_status(1), ' after step ', done
          CALL b200_stop_with(INT(col_status(1)))
       END IF
       IF (wrote) THEN
          ! S8 (mo_grotz.f90:340-398): the device captured the record where the reference calls output()
          CALL b200_check(samsim_b200_get_snapshot(h, snap_sc, snap_ar, 0_C_INT32_T, 1_C_INT32_T), 'snapshot')
          o_T = snap_ar(:,1);  o_psi_s = snap_ar(:,2);  o_thick = snap_ar(:,3);  o_S_bu = snap_ar(:,4)
          o_ray = snap_ar(1:Nlayer-1,5);  o_psi_l = snap_ar(:,6);  o_perm = snap_ar(:,7)
          o_fv = snap_ar(:,8);  o_fh = snap_ar(:,9);  o_psi_g = snap_ar(:,10)
          o_mto = snap_sc(16:18)
          CALL output(Nlayer,o_T,o_psi_s,o_psi_l,o_thick,o_S_bu,o_ray,format_T,format_psi,format_thick,format_snow,     &
               & snap_sc(1),snap_sc(2),snap_sc(3),snap_sc(4),snap_sc(5),snap_sc(6),snap_sc(7),snap_sc(8),snap_sc(9),   &
               & snap_sc(10),snap_sc(11),snap_sc(12),snap_sc(13),snap_sc(14),snap_sc(15),o_perm,format_perm,o_fv,o_fh, &
               & o_psi_g,o_mto,format_melt)
          IF (bgc_flag == 2) THEN   ! output_bgc (mo_output.f90:156-188): the device captured bgc_bu / bgc_br per tracer
             DO k = 1, N_bgc
                WRITE(2*k+400,format_bgc) snap_ar(:,9+2*k)
                WRITE(2*k+401,format_bgc) snap_ar(:,10+2*k)
             END DO
          END IF
          WRITE(*,'(A10,I3,A15,F6.3)') 'progress: ', INT(100._wp*snap_sc(19)/time_total), '%,  thickness: ', snap_sc(9)
       END IF
    END DO

    !---- unchanged reference epilogue (mo_grotz.f90:840-876) ----------------------------------------------------
    CALL b200_pull_mo_data(h)
    CALL samsim_b200_destroy(h)
    WRITE(*,*)'Run completed, total ice thickness at end of run:',SUM(thick(1:N_active-1)),' melt_err= ', melt_err
    CLOSE(30); CLOSE(31); CLOSE(32); CLOSE(33); CLOSE(34); CLOSE(35)
    CLOSE(40); CLOSE(41); CLOSE(42); CLOSE(43); CLOSE(44); CLOSE(45); CLOSE(46); CLOSE(47); CLOSE(48); CLOSE(49)
    CLOSE(50); CLOSE(66)
    IF (bgc_flag == 2) THEN                 ! mo_grotz.f90:850-855
       DO k = 1, N_bgc
          CLOSE(2*k+400); CLOSE(2*k+401)
       END DO
    END IF
    CALL sub_deallocate
  END SUBROUTINE grotz_batch

END MODULE mo_grotz
