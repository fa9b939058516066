c     ****************************************************************
c     *                                                              *
c     *               subroutine w3d_dump_block                      *
c     *                                                              *
c     *   DEBUG HOOK FOR AN ifort BUILD OF THE UNMODIFIED REFERENCE. *
c     *   Insert                                                     *
c     *       call w3d_dump_block( blk, step, iter, local_work )     *
c     *   just before the final return of do_nleps_block             *
c     *   (src/drive_eps_sig_internal_forces.f:399-632).  Every call *
c     *   appends one record to the stream file w3d_dump.bin : the   *
c     *   block's inputs (ue, due, ce_0, history at n, urcs at n)    *
c     *   and its results (ddtse, urcs at n+1, history at n+1, rot,  *
c     *   element forces).  tools/compare_fortran_dump.py replays    *
c     *   each record through the CUDA engine and the CPU            *
c     *   restatement and reports the relative differences: the      *
c     *   function-level 1e-10 pin that cannot be produced without   *
c     *   the Fortran binary.  Arrays are written in their native    *
c     *   (mxvl, ...) shapes; the header gives the sizes.            *
c     *   NOT COMPILED here: see fortran/README.md.                  *
c     *                                                              *
c     ****************************************************************
c
      subroutine w3d_dump_block( blk, step, iter, local_work )
      implicit none
      include 'param_def'
      include 'include_sig_up'
      integer, intent(in) :: blk, step, iter
      integer, save :: unit = -1
      integer :: span, ngp, hist_size, totdof
c
!$OMP CRITICAL (w3d_dump)
      if( unit .lt. 0 ) then
        open( newunit=unit, file='w3d_dump.bin', access='stream',
     &        form='unformatted', status='replace' )
      end if
      span      = local_work%span
      ngp       = local_work%num_int_points
      hist_size = local_work%hist_size_for_blk
      totdof    = local_work%totdof
c             header: magic, sizes, flags
      write(unit) 87033001, blk, step, iter, span, ngp, hist_size,
     &            totdof, mxvl, local_work%mat_type,
     &            merge(1,0,local_work%geo_non_flg),
     &            merge(1,0,local_work%bbar_flg), local_work%felem
c             inputs
      write(unit) local_work%ce_0(1:mxvl,1:totdof)
      write(unit) local_work%ue(1:mxvl,1:totdof)
      write(unit) local_work%due(1:mxvl,1:totdof)
      write(unit) local_work%elem_hist(1:span,1:hist_size,1:ngp)
      write(unit) local_work%urcs_blk_n(1:mxvl,1:nstrs,1:ngp)
      write(unit) local_work%strain_n(1:mxvl,1:nstr,1:ngp)
c             results
      write(unit) local_work%ddtse(1:mxvl,1:nstr,1:ngp)
      write(unit) local_work%urcs_blk_n1(1:mxvl,1:nstrs,1:ngp)
      write(unit) local_work%elem_hist1(1:span,1:hist_size,1:ngp)
      write(unit) local_work%rot_blk_n1(1:mxvl,1:9,1:ngp)
      flush( unit )
!$OMP END CRITICAL (w3d_dump)
      return
      end
