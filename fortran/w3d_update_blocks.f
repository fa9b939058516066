c     ****************************************************************
c     *                                                              *
c     *   block part of update / adaptive_save / adaptive_reset      *
c     *   (B200): the n+1 -> n and n -> n+1 copies of the blocked    *
c     *   state (src/update.f:75-93, src/mnralg.f:1258-1420) happen  *
c     *   on the device where the state lives.                       *
c     *   NOT COMPILED here: see fortran/README.md.                  *
c     *                                                              *
c     ****************************************************************
c
      subroutine update_copy_blocks_w3d
      use w3d_engine_mod
      implicit none
      integer(c_int32_t) :: ierr
c             history, eps, urcs : n+1 -> n   (update.f:75-93)
      ierr = w3d_update( w3d_handle )
      call w3d_check( ierr, 'w3d_update' )
      return
      end
c
      subroutine adaptive_save
      use w3d_engine_mod
      implicit none
      integer(c_int32_t) :: ierr
c             rot_n1 -> rot_n   (mnralg.f:1258-1310)
      ierr = w3d_adaptive_save( w3d_handle )
      call w3d_check( ierr, 'w3d_adaptive_save' )
      return
      end
c
      subroutine adaptive_reset
      use w3d_engine_mod
      implicit none
      integer(c_int32_t) :: ierr
c             history, rot, eps, urcs : n -> n+1  (mnralg.f:1335-1401)
      ierr = w3d_adaptive_reset( w3d_handle )
      call w3d_check( ierr, 'w3d_adaptive_reset' )
      return
      end
c
c             readers of blocked state outside the path (output ou*.f,
c             oustates.f, J-integrals di*.f, store.f) refresh the host
c             copy of one block first:
c
      subroutine w3d_refresh_block( blk )
      use elem_block_data, only : history_blocks, history1_blocks,
     &       eps_n_blocks, eps_n1_blocks, urcs_n_blocks,
     &       urcs_n1_blocks, rot_n_blocks, rot_n1_blocks, cep_blocks
      use w3d_engine_mod
      implicit none
      integer, intent(in) :: blk
      integer(c_int32_t) :: ierr, b
      b = int( blk, c_int32_t )
      ierr = w3d_pull_state( w3d_handle, w3d_hist_n, b,
     &                       history_blocks(blk)%ptr )
      ierr = ierr + w3d_pull_state( w3d_handle, w3d_hist_n1, b,
     &                       history1_blocks(blk)%ptr )
      ierr = ierr + w3d_pull_state( w3d_handle, w3d_eps_n, b,
     &                       eps_n_blocks(blk)%ptr )
      ierr = ierr + w3d_pull_state( w3d_handle, w3d_eps_n1, b,
     &                       eps_n1_blocks(blk)%ptr )
      ierr = ierr + w3d_pull_state( w3d_handle, w3d_urcs_n, b,
     &                       urcs_n_blocks(blk)%ptr )
      ierr = ierr + w3d_pull_state( w3d_handle, w3d_urcs_n1, b,
     &                       urcs_n1_blocks(blk)%ptr )
      ierr = ierr + w3d_pull_state( w3d_handle, w3d_rot_n, b,
     &                       rot_n_blocks(blk)%ptr )
      ierr = ierr + w3d_pull_state( w3d_handle, w3d_rot_n1, b,
     &                       rot_n1_blocks(blk)%ptr )
      ierr = ierr + w3d_pull_state( w3d_handle, w3d_cep, b,
     &                       cep_blocks(blk)%vector )
      call w3d_check( ierr, 'w3d_pull_state' )
      return
      end
