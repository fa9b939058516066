c     ****************************************************************
c     *                                                              *
c     *                   module w3d_engine_mod                      *
c     *                                                              *
c     *   ISO_C_BINDING interface to the B200 element-block engine   *
c     *   (include/w3d_engine.h, ABI version 3).  One interface per  *
c     *   exported entry point; derived types mirror the C structs   *
c     *   field for field.  NOT COMPILED in the development image    *
c     *   (no Fortran compiler there): see fortran/README.md.        *
c     *                                                              *
c     ****************************************************************
c
      module w3d_engine_mod
      use iso_c_binding
      implicit none
c
      integer(c_int32_t), parameter :: w3d_abi_version = 3
c                 selectors of w3d_push_state / w3d_pull_state
      integer(c_int32_t), parameter :: w3d_hist_n = 0, w3d_hist_n1 = 1,
     &   w3d_eps_n = 2, w3d_eps_n1 = 3, w3d_urcs_n = 4,
     &   w3d_urcs_n1 = 5, w3d_rot_n = 6, w3d_rot_n1 = 7, w3d_cep = 8,
     &   w3d_estiff = 9, w3d_elem_vol = 10
c
      type, bind(c) :: w3d_engine_config
        integer(c_int32_t) :: abi_version
        integer(c_int32_t) :: mxvl, mxelpr, mxmtpr
        integer(c_int32_t) :: nelblk, noelem, nonode, nodof
        integer(c_int32_t) :: myid, numprocs
        type(c_ptr) :: elblks          ! elblks(0:3,nelblk)   inelbk.f:47-52
        type(c_ptr) :: props           ! props(mxelpr,noelem) mod_main.f:41-46
        type(c_ptr) :: matprp          ! matprp(mxmtpr,nummat)
        integer(c_int32_t) :: nmat
        type(c_ptr) :: c               ! c(nodof)
        type(c_ptr) :: edest           ! packed edest_blocks(blk)%ptr(totdof,span)
        type(c_ptr) :: cdest           ! packed cdest_blocks
        type(c_ptr) :: history_blk_list
        real(c_double) :: eps_bbar
        integer(c_int32_t) :: qbar_flag
        real(c_double) :: beta_fact
        integer(c_int32_t) :: adaptive_flag
        integer(c_int32_t) :: device
c                 crystal plasticity (mm10); c_null_ptr / 0 without CP
        type(c_ptr) :: crystal_types   ! n x 344 doubles, c_array(cnum)
        integer(c_int32_t) :: n_crystal_types
        integer(c_int32_t) :: cp_ncrystals      ! imatprp(101,matnum)
        type(c_ptr) :: elem_cnum       ! (cp_ncrystals,noelem)
        type(c_ptr) :: elem_angles     ! (3,cp_ncrystals,noelem)
        integer(c_int32_t) :: cp_angle_type     ! imatprp(103)
        integer(c_int32_t) :: asymmetric_assembly
      end type
c
      type, bind(c) :: w3d_scalars
        real(c_double) :: internal_energy, plastic_work, sum_ifv
        integer(c_int64_t) :: num_term_ifv
        integer(c_int32_t) :: material_cut_step, reserved
      end type
c
      type, bind(c) :: w3d_residual_loads
        type(c_ptr) :: cstmap, pbar, load, mdiag, velocity, accel
        type(c_ptr) :: contact_force, d_lagrange_forces
      end type
      type, bind(c) :: w3d_residual_norms
        real(c_double) :: norm_res, mgload, sum_loads
        integer(c_int64_t) :: num_term_loads
        real(c_double) :: res_max
        integer(c_int32_t) :: res_max_dof
      end type
c
c                 the one handle of this rank (one GPU per MPI rank)
      type(c_ptr), save :: w3d_handle = c_null_ptr
c                 packed copies of edest_blocks / cdest_blocks kept for
c                 w3d_assemble_setup
      integer(c_int32_t), allocatable, save, target ::
     &   w3d_edest_packed(:), w3d_cdest_packed(:)
c
      interface
        integer(c_int32_t) function w3d_engine_create( cfg, handle )
     &     bind(c, name='w3d_engine_create')
          import
          type(w3d_engine_config), intent(in) :: cfg
          type(c_ptr), intent(out) :: handle
        end function
        integer(c_int32_t) function w3d_engine_destroy( h )
     &     bind(c, name='w3d_engine_destroy')
          import
          type(c_ptr), value :: h
        end function
        type(c_ptr) function w3d_engine_last_error( h )
     &     bind(c, name='w3d_engine_last_error')
          import
          type(c_ptr), value :: h
        end function
        integer(c_int32_t) function w3d_set_time_step( h, dt )
     &     bind(c, name='w3d_set_time_step')
          import
          type(c_ptr), value :: h
          real(c_double), value :: dt
        end function
        integer(c_int32_t) function w3d_eps_sig_ifv( h, step, iter,
     &     extrapolated_du, u, du, ifv, scal )
     &     bind(c, name='w3d_eps_sig_ifv')
          import
          type(c_ptr), value :: h
          integer(c_int32_t), value :: step, iter, extrapolated_du
          real(c_double), intent(in) :: u(*), du(*)
          real(c_double), intent(out) :: ifv(*)
          type(w3d_scalars), intent(out) :: scal
        end function
        integer(c_int32_t) function w3d_tanstf( h, step, iter, u, du,
     &     estiff ) bind(c, name='w3d_tanstf')
          import
          type(c_ptr), value :: h
          integer(c_int32_t), value :: step, iter
          type(c_ptr), value :: u, du     ! c_null_ptr: resident vectors
          type(c_ptr), value :: estiff    ! c_null_ptr: keep on device
        end function
        integer(c_int32_t) function w3d_update( h )
     &     bind(c, name='w3d_update')
          import
          type(c_ptr), value :: h
        end function
        integer(c_int32_t) function w3d_adaptive_save( h )
     &     bind(c, name='w3d_adaptive_save')
          import
          type(c_ptr), value :: h
        end function
        integer(c_int32_t) function w3d_adaptive_reset( h )
     &     bind(c, name='w3d_adaptive_reset')
          import
          type(c_ptr), value :: h
        end function
        integer(c_int32_t) function w3d_push_state( h, which, blk,
     &     src ) bind(c, name='w3d_push_state')
          import
          type(c_ptr), value :: h
          integer(c_int32_t), value :: which, blk
          real(c_double), intent(in) :: src(*)
        end function
        integer(c_int32_t) function w3d_pull_state( h, which, blk,
     &     dst ) bind(c, name='w3d_pull_state')
          import
          type(c_ptr), value :: h
          integer(c_int32_t), value :: which, blk
          real(c_double), intent(out) :: dst(*)
        end function
        integer(c_int64_t) function w3d_block_offset( h, which, blk )
     &     bind(c, name='w3d_block_offset')
          import
          type(c_ptr), value :: h
          integer(c_int32_t), value :: which, blk
        end function
c
c                 exchange step between MPI ranks (NCCL inside the engine)
        integer(c_int32_t) function w3d_comm_unique_id( id )
     &     bind(c, name='w3d_comm_unique_id')
          import
          integer(c_int8_t), intent(out) :: id(128)
        end function
        integer(c_int32_t) function w3d_comm_init( h, id, rank,
     &     nranks ) bind(c, name='w3d_comm_init')
          import
          type(c_ptr), value :: h
          integer(c_int8_t), intent(in) :: id(128)
          integer(c_int32_t), value :: rank, nranks
        end function
        integer(c_int32_t) function w3d_comm_set_options( h, exchange,
     &     bcast_inputs ) bind(c, name='w3d_comm_set_options')
          import
          type(c_ptr), value :: h
          integer(c_int32_t), value :: exchange, bcast_inputs
        end function
        integer(c_int32_t) function w3d_download_ifv_range( h, first,
     &     n, dst ) bind(c, name='w3d_download_ifv_range')
          import
          type(c_ptr), value :: h
          integer(c_int32_t), value :: first, n
          real(c_double), intent(out) :: dst(*)
        end function
c
c                 assembled equations (assemble_code.f on the device)
        integer(c_int32_t) function w3d_assemble_setup( h, cstmap,
     &     edest, neqns, ncoeff ) bind(c, name='w3d_assemble_setup')
          import
          type(c_ptr), value :: h
          integer(c_int32_t), intent(in) :: cstmap(*), edest(*)
          integer(c_int32_t), intent(out) :: neqns
          integer(c_int64_t), intent(out) :: ncoeff
        end function
        integer(c_int32_t) function w3d_assemble_structure( h,
     &     dof_eqn_map, eqn_node_map, k_ptrs, k_indexes )
     &     bind(c, name='w3d_assemble_structure')
          import
          type(c_ptr), value :: h
          integer(c_int32_t), intent(out) :: dof_eqn_map(*),
     &       eqn_node_map(*), k_ptrs(*), k_indexes(*)
        end function
        integer(c_int32_t) function w3d_assemble_rows( h, first_eqn,
     &     n_eqns ) bind(c, name='w3d_assemble_rows')
          import
          type(c_ptr), value :: h
          integer(c_int32_t), intent(out) :: first_eqn, n_eqns
        end function
        integer(c_int32_t) function w3d_assemble( h, k_diag,
     &     k_coeffs ) bind(c, name='w3d_assemble')
          import
          type(c_ptr), value :: h
          real(c_double), intent(out) :: k_diag(*), k_coeffs(*)
        end function
        integer(c_int32_t) function w3d_assemble_csr( h, nnz, k_ptrs,
     &     k_indexes, k_coeffs ) bind(c, name='w3d_assemble_csr')
          import
          type(c_ptr), value :: h
          integer(c_int64_t), intent(out) :: nnz
          integer(c_int32_t), intent(out) :: k_ptrs(*), k_indexes(*)
          real(c_double), intent(out) :: k_coeffs(*)
        end function
c
c                 structure residual and its norms (upres, cvtest)
        integer(c_int32_t) function w3d_residual_set_loads( h, loads )
     &     bind(c, name='w3d_residual_set_loads')
          import
          type(c_ptr), value :: h
          type(w3d_residual_loads), intent(in) :: loads
        end function
        integer(c_int32_t) function w3d_residual( h, iter, dt, nbeta,
     &     res, norms ) bind(c, name='w3d_residual')
          import
          type(c_ptr), value :: h
          integer(c_int32_t), value :: iter
          real(c_double), value :: dt, nbeta
          real(c_double), intent(out) :: res(*)
          type(w3d_residual_norms), intent(out) :: norms
        end function
      end interface
c
      contains
c
c                 die with the engine's message (all ranks)
      subroutine w3d_check( ierr, where )
      integer(c_int32_t), intent(in) :: ierr
      character(len=*), intent(in) :: where
      character(kind=c_char), pointer :: msg(:)
      type(c_ptr) :: p
      integer :: k
      if( ierr .eq. 0 ) return
      p = w3d_engine_last_error( w3d_handle )
      write(*,9000) where, ierr
      if( c_associated( p ) ) then
        call c_f_pointer( p, msg, [512] )
        do k = 1, 512
          if( msg(k) .eq. c_null_char ) exit
          write(*,'(a1)',advance='no') msg(k)
        end do
        write(*,*)
      end if
      call die_abort
 9000 format(/,'>>>>> FATAL ERROR: w3d engine, ',a,', code ',i3)
      end subroutine
c
      end module w3d_engine_mod
