c     ****************************************************************
c     *                                                              *
c     *                 subroutine tanstf  (B200)                    *
c     *                                                              *
c     *   drop-in replacement of the body of src/tanstf.f:17-95 :    *
c     *   estiff_allocate + the OpenMP loop over do_nlek_block       *
c     *   (dptstf, rktstf, gptns1, ctran1, bdbtgen, kgstiff,         *
c     *   rktstf_do_transpose) become ONE call (w3d_tanstf).         *
c     *                                                              *
c     *   keep_on_device = .true. (use with w3d_assemble: the        *
c     *   sparse equations are formed on the GPU and only k_diag /   *
c     *   k_coeffs cross to the host) or .false.: the element        *
c     *   matrices come back in estiff_blocks(blk)%ptr(utsz,span),   *
c     *   the layout the reference's assembler reads.                *
c     *   NOT COMPILED here: see fortran/README.md.                  *
c     *                                                              *
c     ****************************************************************
c
      subroutine tanstf( first, step, iter )
      use global_data, only : nelblk, elblks, myid
      use elem_block_data, only : estiff_blocks
      use w3d_engine_mod
c
      implicit none
c
      logical, intent(in) :: first
      integer, intent(in) :: step, iter
c
      logical, parameter :: keep_on_device = .false.
      integer :: blk
      integer(c_int32_t) :: ierr
      integer(c_int64_t) :: ntot, off
      double precision, allocatable, target, save :: packed(:)
c
      if( keep_on_device ) then
        ierr = w3d_tanstf( w3d_handle, int(step,c_int32_t),
     &                     int(iter,c_int32_t), c_null_ptr,
     &                     c_null_ptr, c_null_ptr )
        call w3d_check( ierr, 'w3d_tanstf' )
        return
      end if
c
c             one packed buffer for the owned blocks in block order;
c             block blk starts at w3d_block_offset( h, estiff, blk )
c
      ntot = w3d_block_offset( w3d_handle, w3d_estiff,
     &                         int(nelblk+1,c_int32_t) )
      if( .not. allocated( packed ) ) allocate( packed(ntot) )
      ierr = w3d_tanstf( w3d_handle, int(step,c_int32_t),
     &                   int(iter,c_int32_t), c_null_ptr, c_null_ptr,
     &                   c_loc(packed) )
      call w3d_check( ierr, 'w3d_tanstf' )
c
      call estiff_allocate( 1 )
      do blk = 1, nelblk
        if( elblks(2,blk) .ne. myid ) cycle
        off = w3d_block_offset( w3d_handle, w3d_estiff,
     &                          int(blk,c_int32_t) )
c             (utsz,span) in the engine = (utsz,span) in the reference:
c             no transpose (rktstf_do_transpose is gone)
        estiff_blocks(blk)%ptr = reshape(
     &     packed(off+1:off+size(estiff_blocks(blk)%ptr)),
     &     shape(estiff_blocks(blk)%ptr) )
      end do
c
      return
      end
