c     ****************************************************************
c     *                                                              *
c     *               subroutine w3d_engine_start  (B200)            *
c     *                                                              *
c     *   called once per analysis after history_cep_init,           *
c     *   edest_init and cdest_init (src/reopen.f:1015-1408): hands  *
c     *   the model tables to the engine and creates this rank's     *
c     *   handle.  Under MPI every rank calls it (one GPU per rank); *
c     *   root creates the NCCL id, MPI broadcasts it, every rank    *
c     *   joins the engine's communicator -- from then on the        *
c     *   wmpi_reduce_vec(ifv) of the stress pass is done by the     *
c     *   engine.  On a restart follow with w3d_push_state for every *
c     *   owned block and array kind.                                *
c     *   NOT COMPILED here: see fortran/README.md.                  *
c     *                                                              *
c     ****************************************************************
c
      subroutine w3d_engine_start
      use global_data, only : nelblk, noelem, nonode, nodof, myid,
     &       numprocs, elblks, props, matprp, nummat, eps_bbar,
     &       qbar_flag, beta_fact, adaptive_flag, use_mpi, c, mxvl,
     &       mxelpr, mxmtpr
      use main_data, only : asymmetric_assembly
      use elem_block_data, only : edest_blocks, cdest_blocks,
     &                            history_blk_list
      use w3d_engine_mod
      implicit none
      include 'mpif.h'
c
      type(w3d_engine_config) :: cfg
      integer :: blk, span, felem, totdof, i, j, ierr_mpi
      integer(c_int64_t) :: ntab, pos
      integer(c_int32_t) :: ierr
      integer(c_int8_t) :: nccl_id(128)
      integer(c_int32_t), allocatable, target, save :: hbl(:)
c
c             pack the per-block (totdof,span) index tables in block
c             order -- the layout edest_blocks(blk)%ptr already has
c
      ntab = 0
      do blk = 1, nelblk
        ntab = ntab + size( edest_blocks(blk)%ptr )
      end do
      allocate( w3d_edest_packed(ntab), w3d_cdest_packed(ntab) )
      pos = 0
      do blk = 1, nelblk
        span   = elblks(0,blk)
        totdof = size( edest_blocks(blk)%ptr, 1 )
        do i = 1, span
          do j = 1, totdof
            pos = pos + 1
            w3d_edest_packed(pos) = edest_blocks(blk)%ptr(j,i)
            w3d_cdest_packed(pos) = cdest_blocks(blk)%ptr(j,i)
          end do
        end do
      end do
      allocate( hbl(nelblk) )
      hbl(1:nelblk) = history_blk_list(1:nelblk)
c
      cfg%abi_version = w3d_abi_version
      cfg%mxvl = mxvl;  cfg%mxelpr = mxelpr;  cfg%mxmtpr = mxmtpr
      cfg%nelblk = nelblk;  cfg%noelem = noelem
      cfg%nonode = nonode;  cfg%nodof  = nodof
      cfg%myid = myid;  cfg%numprocs = numprocs
      cfg%elblks = c_loc( elblks )
      cfg%props  = c_loc( props )
      cfg%matprp = c_loc( matprp )
      cfg%nmat   = nummat
      cfg%c      = c_loc( c )
      cfg%edest  = c_loc( w3d_edest_packed )
      cfg%cdest  = c_loc( w3d_cdest_packed )
      cfg%history_blk_list = c_loc( hbl )
      cfg%eps_bbar = eps_bbar
      cfg%qbar_flag = merge( 1, 0, qbar_flag )
      cfg%beta_fact = beta_fact
      cfg%adaptive_flag = merge( 1, 0, adaptive_flag )
      cfg%device = -1           ! the device this rank selected
      cfg%crystal_types = c_null_ptr   ! see w3d_engine.h for the
      cfg%n_crystal_types = 0          ! 344-double record built from
      cfg%cp_ncrystals = 0             ! c_array(cnum) when a block
      cfg%elem_cnum = c_null_ptr       ! uses material model 10
      cfg%elem_angles = c_null_ptr
      cfg%cp_angle_type = 1
      cfg%asymmetric_assembly = merge( 1, 0, asymmetric_assembly )
c
      ierr = w3d_engine_create( cfg, w3d_handle )
      call w3d_check( ierr, 'w3d_engine_create' )
c
      if( use_mpi .and. numprocs .gt. 1 ) then
        if( myid .eq. 0 ) then
          ierr = w3d_comm_unique_id( nccl_id )
          call w3d_check( ierr, 'w3d_comm_unique_id' )
        end if
        call MPI_BCAST( nccl_id, 128, MPI_BYTE, 0, MPI_COMM_WORLD,
     &                  ierr_mpi )
        ierr = w3d_comm_init( w3d_handle, nccl_id,
     &                        int(myid,c_int32_t),
     &                        int(numprocs,c_int32_t) )
        call w3d_check( ierr, 'w3d_comm_init' )
c             exchange = 0: every rank (hence root) ends with the full
c             ifv, the reference's wmpi_reduce_vec result
        ierr = w3d_comm_set_options( w3d_handle, 0_c_int32_t,
     &                               0_c_int32_t )
        call w3d_check( ierr, 'w3d_comm_set_options' )
      end if
c
      return
      end
