c     ****************************************************************
c     *                                                              *
c     *      subroutine drive_eps_sig_internal_forces  (B200)        *
c     *                                                              *
c     *   drop-in replacement of the body of                         *
c     *   src/drive_eps_sig_internal_forces.f:20-349 : strain        *
c     *   recovery, stress update and internal forces of every       *
c     *   element block this rank owns, in ONE call into the engine  *
c     *   (include/w3d_engine.h: w3d_eps_sig_ifv).  What disappears: *
c     *   allocate_ifv, the OpenMP loop over do_nleps_block, the     *
c     *   block_energies / step_cut_flags reductions (:210-226), the *
c     *   atomic addifv loop (:272-300), and -- after w3d_comm_init  *
c     *   -- wmpi_redlog / wmpi_reduce_vec / wmpi_redint             *
c     *   (:248-250, :338-340): the engine exchanges over NCCL.      *
c     *                                                              *
c     *   Same arguments, same module variables on exit (ifv,        *
c     *   internal_energy, plastic_work, sum_ifv, num_term_ifv,      *
c     *   material_cut_step).  NOT COMPILED here: fortran/README.md. *
c     *                                                              *
c     ****************************************************************
c
      subroutine drive_eps_sig_internal_forces( step, iter,
     &                                          material_cut_step )
      use global_data, only : out, ifv, internal_energy, nodof,
     &                        num_term_ifv, plastic_work, sum_ifv,
     &                        u, du, dt
      use main_data, only : extrapolated_du
      use w3d_engine_mod
      use constants
c
      implicit none
c
      integer, intent(in) :: step, iter
      logical, intent(inout) :: material_cut_step
c
      type(w3d_scalars) :: scal
      integer(c_int32_t) :: ierr, iextrap
c
c             rate-dependent models (mm10) read the step's time
c             increment; cheap, so pass it every call
c
      ierr = w3d_set_time_step( w3d_handle, dt )
      call w3d_check( ierr, 'w3d_set_time_step' )
c
c             u, du: displacements at n and the current estimate of
c             the increment over the step (global_data).  ifv comes
c             back as this rank's contribution, or -- with the
c             engine's communicator -- complete on every dof this
c             rank's blocks touch (exchange = 1) / on all dofs
c             (exchange = 0), already scaled by beta_fact and zeroed
c             when a step cut is requested (:253-263)
c
      iextrap = 0
      if( extrapolated_du ) iextrap = 1
      ierr = w3d_eps_sig_ifv( w3d_handle, int(step,c_int32_t),
     &                        int(iter,c_int32_t), iextrap, u, du,
     &                        ifv, scal )
      call w3d_check( ierr, 'w3d_eps_sig_ifv' )
c
      material_cut_step = scal%material_cut_step .ne. 0
      internal_energy   = scal%internal_energy
      plastic_work      = scal%plastic_work
      sum_ifv           = scal%sum_ifv
      num_term_ifv      = int( scal%num_term_ifv )
      if( material_cut_step ) then
        sum_ifv      = zero
        num_term_ifv = 0
      end if
c
      return
      end
