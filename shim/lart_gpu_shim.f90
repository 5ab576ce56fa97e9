!------------------------------------------------------------------------------
! lart_gpu_shim.f90 -- ISO_C_BINDING shim between LaRT's Fortran host and the
! B200 engine (include/lart_gpu.h, liblart_gpu.so).
!
! UNTESTED: no Fortran compiler exists in the build image or on the GPU box, so
! this file has never been compiled.  It is deliberately thin: it marshals the
! host's module state (grid, par, line, observer(:), scatt_mat, allph) into the
! POD structs of lart_gpu.h and calls four entry points.  The same sequence is
! exercised, through the same C ABI, by lart_b200/host.py (ctypes) in the tests.
!
! It supplies one more implementation of the `run_sim` interface
! (src/define.f90:832-838):
!
!     subroutine run_gpu(grid)      ! bind with:  run_simulation => run_gpu
!
! to be selected in setup_procedure (src/setup.f90:1001-1005) -- see
! INTEGRATION.md for the two-line host patch.  After it returns, main.f90:46-52
! (output_reduce, output_normalize, write_output) run unchanged: the tallies are
! ADDED into the host's own arrays as raw weighted sums, exactly what
! run_equal_number leaves behind.
!
! MPI: one rank per GPU (device = node-local rank modulo the GPU count).  Each
! rank runs photon ids rank+1, rank+1+nproc, ... (src/run_simulation_mod.f90:150);
! the ranks' tallies are summed on the device by ONE NCCL reduce inside the C ABI
! (lart_gpu_reduce; the communicator id travels by MPI_BCAST) before they are
! added into the host arrays, so the host's plane-by-plane MPI reduce of 160-400 MB
! arrays (src/memory_mod_mpi.f90:366-458) only ever sums zeros onto rank 0's result.
!------------------------------------------------------------------------------
module lart_gpu_shim
  use, intrinsic :: iso_c_binding
  use define
  use clump_mod, only: N_clumps, sphere_R, cl_Dfreq_ref, cl_x, cl_y, cl_z, cl_vx, cl_vy, cl_vz, cl_radius, cl_rhokap, &
                       cl_rhokapD, cl_voigt_a, cl_Dfreq, cgx, cgy, cgz, cg_xmin, cg_ymin, cg_zmin, cg_dx, cg_dy, cg_dz, &
                       cg_start, cg_list, has_overlap
  use octree_mod, only: amr_grid
  implicit none
  private
  public :: run_gpu

  !--- mirrors of the structs in include/lart_gpu.h (same member order) -------
  type, bind(C) :: c_lart_grid
     integer(c_int32_t) :: nx, ny, nz, nxfreq
     real(c_double) :: xmin, ymin, zmin, xmax, ymax, zmax, dx, dy, dz
     real(c_double) :: Dfreq_ref, xfreq_min, xfreq_max, dxfreq, xcrit, xcrit2, rmax
     integer(c_int32_t) :: i0, j0, k0, pad_
     type(c_ptr) :: xface, yface, zface, rhokap, voigt_a, Dfreq, vfx, vfy, vfz, rhokapD
     type(c_ptr) :: mask
     integer(c_int32_t) :: geometry_JPa, nr
     type(c_ptr) :: ind_sph, ind_cyl
  end type
  type, bind(C) :: c_lart_params
     integer(c_int64_t) :: nphotons, seed
     real(c_double) :: xfreq0, xs_point, ys_point, zs_point, source_rmax, DGR, albedo, hgg
     real(c_double) :: voigt_a0, Dfreq0, gaussian_sigma_x, mu_min, dmu
     integer(c_int32_t) :: nmu, spectral_type, source_geometry, comoving_source, recoil
     integer(c_int32_t) :: core_skip, core_skip_global, use_stokes, use_reduced_wgt
     integer(c_int32_t) :: save_Jin, save_Jabs, save_Jmu, save_peeloff, save_peeloff_2D, save_peeloff_3D
     integer(c_int32_t) :: save_direc0, save_all_photons, xyz_symmetry, xy_symmetry, use_clump_medium, xy_periodic, nobs
     integer(c_int32_t) :: use_amr_grid
     integer(c_int32_t) :: atmosphere, calc_J, calc_P, calc_Pnew
     real(c_double) :: Omega
  end type
  type, bind(C) :: c_lart_line
     integer(c_int32_t) :: line_type, pad_
     real(c_double) :: E1, E2, E3, g_recoil0, DnuHK_Hz, cross0
  end type
  type, bind(C) :: c_lart_observer
     real(c_double) :: x, y, z, rmatrix(9), dxim, dyim
     integer(c_int32_t) :: nxim, nyim
  end type
  type, bind(C) :: c_lart_scatt_mat
     integer(c_int32_t) :: nPDF, pad_
     type(c_ptr) :: coss, S11, S12, S33, S34, phase_PDF, alias
  end type
  type, bind(C) :: c_lart_clumps            ! src/clump_mod.f90:30-118
     integer(c_int64_t) :: n
     real(c_double)     :: sphere_R, Dfreq_ref
     type(c_ptr)        :: x, y, z, vx, vy, vz, radius, rhokap, rhokapD, voigt_a, Dfreq
     integer(c_int32_t) :: cgx, cgy, cgz, has_overlap
     real(c_double)     :: cg_xmin, cg_ymin, cg_zmin, cg_dx, cg_dy, cg_dz
     type(c_ptr)        :: cg_start, cg_list
  end type
  type, bind(C) :: c_lart_amr               ! amr_grid_type, src/octree_mod.f90:19-138
     integer(c_int32_t) :: ncells, nleaf
     type(c_ptr)        :: children, ileaf, icell_of_leaf, neighbor
     type(c_ptr)        :: cx, cy, cz, ch
     type(c_ptr)        :: rhokap, voigt_a, Dfreq, vfx, vfy, vfz, rhokapD
  end type
  type, bind(C) :: c_lart_config
     type(c_lart_grid)      :: grid
     type(c_lart_params)    :: par
     type(c_lart_line)      :: line
     type(c_lart_scatt_mat) :: scatt_mat
     type(c_lart_clumps)    :: clumps
     type(c_lart_amr)       :: amr
     type(c_ptr)            :: observers
     integer(c_int32_t)     :: device, pool_slots, quantum, flags, streams, ray_budget, max_events, pad_
  end type
  type, bind(C) :: c_lart_observer_out
     type(c_ptr) :: scatt, direc, direc0, I, Q, U, V
     type(c_ptr) :: scatt_2D, direc_2D, direc0_2D, I_2D, Q_2D, U_2D, V_2D
  end type
  type, bind(C) :: c_lart_allph_out
     type(c_ptr) :: rp0, rp, xfreq1, xfreq2, nscatt_gas, nscatt_dust, I, Q, U, V
  end type
  type, bind(C) :: c_lart_counters
     real(c_double) :: n_photons_done, n_scatter, n_cellsteps, n_peel, n_rng, n_reject_iter, n_peel_bound, n_cellsteps_bound
  end type
  type, bind(C) :: c_lart_tallies
     type(c_ptr) :: Jout, Jin, Jabs, Jmu, obs
     type(c_lart_allph_out) :: allph
     real(c_double) :: nscatt_gas, nscatt_dust
     type(c_lart_counters) :: counters
     type(c_ptr) :: Jabs2, J, Pa, Pnew
  end type

  interface
     integer(c_int) function lart_gpu_create(cfg, handle) bind(C, name='lart_gpu_create')
       import :: c_int, c_ptr, c_lart_config
       type(c_lart_config), intent(in) :: cfg
       type(c_ptr), intent(out) :: handle
     end function
     integer(c_int) function lart_gpu_run(handle, first_id, count, stride) bind(C, name='lart_gpu_run')
       import :: c_int, c_ptr, c_int64_t
       type(c_ptr), value :: handle
       integer(c_int64_t), value :: first_id, count, stride
     end function
     integer(c_int) function lart_gpu_fetch(handle, tallies) bind(C, name='lart_gpu_fetch')
       import :: c_int, c_ptr, c_lart_tallies
       type(c_ptr), value :: handle
       type(c_lart_tallies), intent(inout) :: tallies
     end function
     integer(c_int) function lart_gpu_destroy(handle) bind(C, name='lart_gpu_destroy')
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
     end function
     type(c_ptr) function lart_gpu_last_error() bind(C, name='lart_gpu_last_error')
       import :: c_ptr
     end function
     integer(c_int) function lart_gpu_comm_unique_id(id128) bind(C, name='lart_gpu_comm_unique_id')
       import :: c_int, c_char
       character(kind=c_char), intent(out) :: id128(128)
     end function
     integer(c_int) function lart_gpu_comm_init(device, nranks, rank, id128) bind(C, name='lart_gpu_comm_init')
       import :: c_int, c_int32_t, c_char
       integer(c_int32_t), value :: device, nranks, rank
       character(kind=c_char), intent(in) :: id128(128)
     end function
     integer(c_int) function lart_gpu_reduce(handle, root) bind(C, name='lart_gpu_reduce')
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: handle
       integer(c_int32_t), value :: root
     end function
  end interface

contains

  pure function l2i(flag) result(i)
    logical, intent(in) :: flag
    integer(c_int32_t) :: i
    i = merge(1_c_int32_t, 0_c_int32_t, flag)
  end function

  function ploc(arr) result(p)      ! address of a (possibly unassociated) pointer array
    real(kind=wp), pointer, intent(in) :: arr(..)
    type(c_ptr) :: p
    p = c_null_ptr
    if (associated(arr)) p = c_loc(arr)
  end function

  subroutine check(ierr_c)           ! reference convention: print, then MPI_ABORT (src/setup.f90:132-135)
    use mpi
    integer(c_int), intent(in) :: ierr_c
    character(kind=c_char), pointer :: msg(:)
    integer :: ierr, n
    if (ierr_c == 0) return
    call c_f_pointer(lart_gpu_last_error(), msg, [512])
    n = 1
    do while (n < 512 .and. msg(n) /= c_null_char)
       n = n + 1
    end do
    write(*,'(2a)') 'ERROR (lart_gpu): ', transfer(msg(1:n-1), repeat(' ', n-1))
    call MPI_ABORT(MPI_COMM_WORLD, 1, ierr)
  end subroutine

  !--- one more implementation of run_sim (src/define.f90:832-838) -------------
  subroutine run_gpu(grid)
    use mpi
    type(grid_type), intent(inout) :: grid
    type(c_lart_config),  target :: cfg
    type(c_lart_tallies), target :: tal
    type(c_lart_observer),     allocatable, target :: cobs(:)
    type(c_lart_observer_out), allocatable, target :: oout(:)
    type(c_ptr) :: handle
    integer(c_int64_t) :: first_id, count, stride
    integer :: k, ngpu_per_node, ierr
    character(kind=c_char), save :: nccl_id(128)
    logical, save :: comm_ready = .false.

    !--- inputs that bind other run loops than the one behind lart_gpu_run (src/setup.f90:905-990)
    if (par%nside > 0 .or. line%line_type /= 1) then
       write(*,'(a)') 'ERROR (lart_gpu): HEALPix observers and lines other than line_type 1 are not on the GPU path.'
       call MPI_ABORT(MPI_COMM_WORLD, 1, k)
    endif

    !--- grid_type scalars and arrays (src/define.f90:117-148); arrays stay Fortran-owned
    cfg%grid%nx = grid%nx; cfg%grid%ny = grid%ny; cfg%grid%nz = grid%nz; cfg%grid%nxfreq = grid%nxfreq
    cfg%grid%xmin = grid%xmin; cfg%grid%ymin = grid%ymin; cfg%grid%zmin = grid%zmin
    cfg%grid%xmax = grid%xmax; cfg%grid%ymax = grid%ymax; cfg%grid%zmax = grid%zmax
    cfg%grid%dx = grid%dx; cfg%grid%dy = grid%dy; cfg%grid%dz = grid%dz
    cfg%grid%Dfreq_ref = grid%Dfreq_ref; cfg%grid%xfreq_min = grid%xfreq_min; cfg%grid%xfreq_max = grid%xfreq_max
    cfg%grid%dxfreq = grid%dxfreq; cfg%grid%xcrit = grid%xcrit; cfg%grid%xcrit2 = grid%xcrit2; cfg%grid%rmax = par%rmax
    cfg%grid%i0 = grid%i0; cfg%grid%j0 = grid%j0; cfg%grid%k0 = grid%k0; cfg%grid%pad_ = 0
    cfg%grid%xface = ploc(grid%xface); cfg%grid%yface = ploc(grid%yface); cfg%grid%zface = ploc(grid%zface)
    cfg%grid%rhokap = ploc(grid%rhokap); cfg%grid%voigt_a = ploc(grid%voigt_a); cfg%grid%Dfreq = ploc(grid%Dfreq)
    cfg%grid%vfx = ploc(grid%vfx); cfg%grid%vfy = ploc(grid%vfy); cfg%grid%vfz = ploc(grid%vfz)
    cfg%grid%rhokapD = ploc(grid%rhokapD)
    !--- atmosphere mask and the CALCJ / CALCP / CALCPnew index maps (src/define.f90:158, 166-175)
    cfg%grid%mask = c_null_ptr; cfg%grid%ind_sph = c_null_ptr; cfg%grid%ind_cyl = c_null_ptr
    cfg%grid%geometry_JPa = 0; cfg%grid%nr = 0
    if (associated(grid%mask)) cfg%grid%mask = c_loc(grid%mask)
#if defined (CALCJ) || defined (CALCP) || defined (CALCPnew)
    cfg%grid%geometry_JPa = grid%geometry_JPa; cfg%grid%nr = grid%nr
    if (associated(grid%ind_sph)) cfg%grid%ind_sph = c_loc(grid%ind_sph)
    if (associated(grid%ind_cyl)) cfg%grid%ind_cyl = c_loc(grid%ind_cyl)
#endif

    !--- params_type members the path reads (src/define.f90:209-544)
    cfg%par%nphotons = par%nphotons
    cfg%par%seed     = int(par%iseed, c_int64_t)       ! iseed = 0: draw one from /dev/urandom first (random_mt.f90:915-923)
    cfg%par%xfreq0 = par%xfreq0
    cfg%par%xs_point = par%xs_point; cfg%par%ys_point = par%ys_point; cfg%par%zs_point = par%zs_point
    cfg%par%source_rmax = par%source_rmax; cfg%par%DGR = par%DGR; cfg%par%albedo = par%albedo; cfg%par%hgg = par%hgg
    cfg%par%voigt_a0 = par%voigt_a0; cfg%par%Dfreq0 = par%Dfreq0
    if (par%gaussian_FWHM_vel > 0.0_wp) then
       cfg%par%gaussian_sigma_x = par%gaussian_FWHM_vel/2.3548200450309493_wp/vtherm_total(par%temperature)
    else
       cfg%par%gaussian_sigma_x = par%gaussian_sigma_vel/vtherm_total(par%temperature)
    endif
    cfg%par%mu_min = par%mu_min; cfg%par%dmu = par%dmu; cfg%par%nmu = par%nmu
    select case (trim(par%spectral_type))
    case ('voigt');     cfg%par%spectral_type = 1
    case ('voigt0');    cfg%par%spectral_type = 2
    case ('continuum'); cfg%par%spectral_type = 3
    case ('gaussian');  cfg%par%spectral_type = 4
    case default;       cfg%par%spectral_type = 0
    end select
    select case (trim(par%source_geometry))
    case ('uniform');                  cfg%par%source_geometry = 1
    case ('uniform_sphere', 'sphere'); cfg%par%source_geometry = 2
    case ('plane_illumination');       cfg%par%source_geometry = 3
    case default;                      cfg%par%source_geometry = 0
    end select
    cfg%par%atmosphere = 0
    if (trim(par%geometry) == 'plane_atmosphere')     cfg%par%atmosphere = 1
    if (trim(par%geometry) == 'spherical_atmosphere') cfg%par%atmosphere = 2
    cfg%par%Omega = par%Omega          ! already q*Omega*xrange (src/grid_mod_car.f90:348-350)
    cfg%par%calc_J = 0; cfg%par%calc_P = 0; cfg%par%calc_Pnew = 0
#ifdef CALCJ
    cfg%par%calc_J = 1
#endif
#ifdef CALCP
    cfg%par%calc_P = 1
#endif
#ifdef CALCPnew
    cfg%par%calc_Pnew = 1
#endif
    cfg%par%comoving_source = l2i(par%comoving_source); cfg%par%recoil = l2i(par%recoil)
    cfg%par%core_skip = l2i(par%core_skip); cfg%par%core_skip_global = l2i(par%core_skip_global)
    cfg%par%use_stokes = l2i(par%use_stokes); cfg%par%use_reduced_wgt = l2i(par%use_reduced_wgt)
    cfg%par%save_Jin = l2i(par%save_Jin); cfg%par%save_Jabs = l2i(par%save_Jabs); cfg%par%save_Jmu = l2i(par%save_Jmu)
    cfg%par%save_peeloff = l2i(par%save_peeloff); cfg%par%save_peeloff_2D = l2i(par%save_peeloff_2D)
    cfg%par%save_peeloff_3D = l2i(par%save_peeloff_3D); cfg%par%save_direc0 = l2i(par%save_direc0)
    cfg%par%save_all_photons = l2i(par%save_all_photons); cfg%par%xy_periodic = l2i(par%xy_periodic)
    cfg%par%xyz_symmetry = l2i(par%xyz_symmetry); cfg%par%xy_symmetry = l2i(par%xy_symmetry)
    cfg%par%nobs = merge(par%nobs, 0, par%save_peeloff)
    cfg%par%use_clump_medium = l2i(par%use_clump_medium)

    !--- clump population and its CSR grid (src/clump_mod.f90:30-118); `use clump_mod` provides them
    cfg%clumps%n = 0
    if (par%use_clump_medium) then
       cfg%clumps%n = N_clumps; cfg%clumps%sphere_R = sphere_R; cfg%clumps%Dfreq_ref = cl_Dfreq_ref
       cfg%clumps%x = c_loc(cl_x); cfg%clumps%y = c_loc(cl_y); cfg%clumps%z = c_loc(cl_z)
       cfg%clumps%vx = c_loc(cl_vx); cfg%clumps%vy = c_loc(cl_vy); cfg%clumps%vz = c_loc(cl_vz)
       cfg%clumps%radius = c_loc(cl_radius); cfg%clumps%rhokap = c_loc(cl_rhokap)
       cfg%clumps%rhokapD = c_null_ptr
       if (par%DGR > 0.0_wp .and. associated(cl_rhokapD)) cfg%clumps%rhokapD = c_loc(cl_rhokapD)
       cfg%clumps%voigt_a = c_loc(cl_voigt_a); cfg%clumps%Dfreq = c_loc(cl_Dfreq)
       cfg%clumps%cgx = cgx; cfg%clumps%cgy = cgy; cfg%clumps%cgz = cgz; cfg%clumps%has_overlap = l2i(has_overlap)
       cfg%clumps%cg_xmin = cg_xmin; cfg%clumps%cg_ymin = cg_ymin; cfg%clumps%cg_zmin = cg_zmin
       cfg%clumps%cg_dx = cg_dx; cfg%clumps%cg_dy = cg_dy; cfg%clumps%cg_dz = cg_dz
       cfg%clumps%cg_start = c_loc(cg_start); cfg%clumps%cg_list = c_loc(cg_list)
    endif

    !--- octree (src/octree_mod.f90:19-138): `use octree_mod, only: amr_grid`; the box and the frequency grid travel in cfg%grid
    !--- (grid_create_amr copies them into `grid`, src/grid_mod_amr.f90:1017-1058)
    cfg%par%use_amr_grid = l2i(par%use_amr_grid)
    cfg%amr%ncells = 0; cfg%amr%nleaf = 0
    if (par%use_amr_grid) then
       cfg%amr%ncells = amr_grid%ncells; cfg%amr%nleaf = amr_grid%nleaf
       cfg%amr%children = c_loc(amr_grid%children); cfg%amr%ileaf = c_loc(amr_grid%ileaf)
       cfg%amr%icell_of_leaf = c_loc(amr_grid%icell_of_leaf); cfg%amr%neighbor = c_loc(amr_grid%neighbor)
       cfg%amr%cx = c_loc(amr_grid%cx); cfg%amr%cy = c_loc(amr_grid%cy); cfg%amr%cz = c_loc(amr_grid%cz); cfg%amr%ch = c_loc(amr_grid%ch)
       cfg%amr%rhokap = c_loc(amr_grid%rhokap); cfg%amr%voigt_a = c_loc(amr_grid%voigt_a); cfg%amr%Dfreq = c_loc(amr_grid%Dfreq)
       cfg%amr%vfx = c_loc(amr_grid%vfx); cfg%amr%vfy = c_loc(amr_grid%vfy); cfg%amr%vfz = c_loc(amr_grid%vfz)
       cfg%amr%rhokapD = c_null_ptr
       if (par%DGR > 0.0_wp .and. associated(amr_grid%rhokapD)) cfg%amr%rhokapD = c_loc(amr_grid%rhokapD)
       cfg%grid%xmin = amr_grid%xmin; cfg%grid%xmax = amr_grid%xmax; cfg%grid%ymin = amr_grid%ymin; cfg%grid%ymax = amr_grid%ymax
       cfg%grid%zmin = amr_grid%zmin; cfg%grid%zmax = amr_grid%zmax
       cfg%grid%Dfreq_ref = amr_grid%Dfreq_ref; cfg%grid%nxfreq = amr_grid%nxfreq
       cfg%grid%xfreq_min = amr_grid%xfreq_min; cfg%grid%xfreq_max = amr_grid%xfreq_max; cfg%grid%dxfreq = amr_grid%dxfreq
       cfg%grid%xcrit = amr_grid%xcrit; cfg%grid%xcrit2 = amr_grid%xcrit2
       ! tallies of an octree run live in amr_grid%Jout/Jin/Jabs/Jmu: lart_gpu_fetch is pointed at those (see fetch below)
    endif

    !--- line_type (src/define.f90:639-656)
    cfg%line%line_type = line%line_type; cfg%line%pad_ = 0
    cfg%line%E1 = line%E1; cfg%line%E2 = line%E2; cfg%line%E3 = line%E3
    cfg%line%g_recoil0 = line%g_recoil0; cfg%line%DnuHK_Hz = line%DnuHK_Hz; cfg%line%cross0 = line%cross0

    !--- scattering_matrix_type (src/define.f90:616-625), dust + Stokes only
    cfg%scatt_mat%nPDF = 0; cfg%scatt_mat%pad_ = 0
    cfg%scatt_mat%coss = c_null_ptr; cfg%scatt_mat%S11 = c_null_ptr; cfg%scatt_mat%S12 = c_null_ptr
    cfg%scatt_mat%S33 = c_null_ptr;  cfg%scatt_mat%S34 = c_null_ptr
    cfg%scatt_mat%phase_PDF = c_null_ptr; cfg%scatt_mat%alias = c_null_ptr
    if (par%DGR > 0.0_wp .and. par%use_stokes) then
       cfg%scatt_mat%nPDF = scatt_mat%nPDF
       cfg%scatt_mat%coss = c_loc(scatt_mat%coss); cfg%scatt_mat%S11 = c_loc(scatt_mat%S11)
       cfg%scatt_mat%S12  = c_loc(scatt_mat%S12);  cfg%scatt_mat%S33 = c_loc(scatt_mat%S33)
       cfg%scatt_mat%S34  = c_loc(scatt_mat%S34);  cfg%scatt_mat%phase_PDF = c_loc(scatt_mat%phase_PDF)
       cfg%scatt_mat%alias = c_loc(scatt_mat%alias)
    endif

    !--- observers (src/define.f90:547-600) and their output cubes
    allocate(cobs(max(1,cfg%par%nobs)), oout(max(1,cfg%par%nobs)))
    do k = 1, cfg%par%nobs
       cobs(k)%x = observer(k)%x; cobs(k)%y = observer(k)%y; cobs(k)%z = observer(k)%z
       cobs(k)%rmatrix = reshape(observer(k)%rmatrix, [9])      ! Fortran memory order, as lart_gpu.h expects
       cobs(k)%dxim = observer(k)%dxim; cobs(k)%dyim = observer(k)%dyim
       cobs(k)%nxim = observer(k)%nxim; cobs(k)%nyim = observer(k)%nyim
       oout(k)%scatt = ploc(observer(k)%scatt);   oout(k)%direc = ploc(observer(k)%direc)
       oout(k)%direc0 = ploc(observer(k)%direc0)
       oout(k)%I = ploc(observer(k)%I); oout(k)%Q = ploc(observer(k)%Q)
       oout(k)%U = ploc(observer(k)%U); oout(k)%V = ploc(observer(k)%V)
       oout(k)%scatt_2D = ploc(observer(k)%scatt_2D); oout(k)%direc_2D = ploc(observer(k)%direc_2D)
       oout(k)%direc0_2D = ploc(observer(k)%direc0_2D)
       oout(k)%I_2D = ploc(observer(k)%I_2D); oout(k)%Q_2D = ploc(observer(k)%Q_2D)
       oout(k)%U_2D = ploc(observer(k)%U_2D); oout(k)%V_2D = ploc(observer(k)%V_2D)
    enddo
    cfg%observers = c_loc(cobs)

    !--- one rank per GPU
    ngpu_per_node  = 8
    cfg%device     = mod(mpar%h_rank, ngpu_per_node)
    cfg%pool_slots = 0; cfg%quantum = 0; cfg%flags = 0; cfg%streams = 0; cfg%ray_budget = 0; cfg%max_events = 0; cfg%pad_ = 0

    !--- one NCCL communicator per process, its 128-byte id broadcast by MPI (once, like MPI_INIT)
    if (.not. comm_ready .and. mpar%nproc > 1) then
       if (mpar%p_rank == 0) call check(lart_gpu_comm_unique_id(nccl_id))
       call MPI_BCAST(nccl_id, 128, MPI_CHARACTER, 0, MPI_COMM_WORLD, ierr)
       call check(lart_gpu_comm_init(cfg%device, int(mpar%nproc, c_int32_t), int(mpar%p_rank, c_int32_t), nccl_id))
       comm_ready = .true.
    endif

    call check(lart_gpu_create(cfg, handle))

    !--- photon ids of this rank: ip = rank+1, nphotons, nproc (src/run_simulation_mod.f90:150)
    first_id = mpar%p_rank + 1
    stride   = mpar%nproc
    count    = 0
    if (par%nphotons >= first_id) count = (par%nphotons - first_id)/stride + 1
    call check(lart_gpu_run(handle, first_id, count, stride))

    !--- ONE ncclReduce(sum, f64) over the contiguous device tally buffer onto rank 0 replaces reduce_mem's per-array
    !--- MPI_REDUCE (src/memory_mod_mpi.f90:366-458): afterwards only rank 0 holds non-zero device tallies, every rank
    !--- adds its device tallies into its (zero-initialised) host arrays, and the host's own output_reduce still yields the
    !--- same sums — or is skipped altogether (INTEGRATION.md, "output_reduce")
    call check(lart_gpu_reduce(handle, 0_c_int32_t))
    !--- add the raw weighted sums into the host's arrays
    tal%Jout = ploc(grid%Jout); tal%Jin = ploc(grid%Jin); tal%Jabs = ploc(grid%Jabs); tal%Jmu = ploc(grid%Jmu)
    if (par%use_amr_grid) then  ! an octree run tallies into amr_grid (src/octree_mod.f90:84-88)
       tal%Jout = c_null_ptr; tal%Jin = c_null_ptr; tal%Jabs = c_null_ptr; tal%Jmu = c_null_ptr
       if (allocated(amr_grid%Jout)) tal%Jout = c_loc(amr_grid%Jout)
       if (allocated(amr_grid%Jin))  tal%Jin  = c_loc(amr_grid%Jin)
       if (allocated(amr_grid%Jabs)) tal%Jabs = c_loc(amr_grid%Jabs)
       if (allocated(amr_grid%Jmu))  tal%Jmu  = c_loc(amr_grid%Jmu)
    endif
    tal%Jabs2 = ploc(grid%Jabs2); tal%J = c_null_ptr; tal%Pa = c_null_ptr; tal%Pnew = c_null_ptr
#ifdef CALCJ
    select case (grid%geometry_JPa)        ! src/grid_mod_car.f90:1422-1434
    case (3);  tal%J = ploc(grid%J)
    case (2);  tal%J = ploc(grid%J2)
    case default; tal%J = ploc(grid%J1)
    end select
#endif
#ifdef CALCP
    select case (grid%geometry_JPa)        ! :1385-1395
    case (3);  tal%Pa = ploc(grid%Pa)
    case (2);  tal%Pa = ploc(grid%P2)
    case default; tal%Pa = ploc(grid%P1)
    end select
#endif
#ifdef CALCPnew
    select case (grid%geometry_JPa)        ! :1410-1420
    case (3);  tal%Pnew = ploc(grid%Pa_new)
    case (2);  tal%Pnew = ploc(grid%P2_new)
    case default; tal%Pnew = ploc(grid%P1_new)
    end select
#endif
    tal%obs  = c_loc(oout)
    tal%allph%rp0 = c_null_ptr; tal%allph%rp = c_null_ptr; tal%allph%xfreq1 = c_null_ptr; tal%allph%xfreq2 = c_null_ptr
    tal%allph%nscatt_gas = c_null_ptr; tal%allph%nscatt_dust = c_null_ptr
    tal%allph%I = c_null_ptr; tal%allph%Q = c_null_ptr; tal%allph%U = c_null_ptr; tal%allph%V = c_null_ptr
    if (par%save_all_photons) then
       ! allph arrays are node-shared windows (grid_mod_car.f90:1130-1146): only one rank per node may add;
       ! with several ranks per node use private copies and sum them as output_sum_rect.f90:129-146 does.
       tal%allph%rp0 = ploc(allph%rp0); tal%allph%rp = ploc(allph%rp)
       tal%allph%xfreq1 = ploc(allph%xfreq1); tal%allph%xfreq2 = ploc(allph%xfreq2)
       tal%allph%nscatt_gas = ploc(allph%nscatt_gas); tal%allph%nscatt_dust = ploc(allph%nscatt_dust)
       tal%allph%I = ploc(allph%I); tal%allph%Q = ploc(allph%Q); tal%allph%U = ploc(allph%U); tal%allph%V = ploc(allph%V)
    endif
    tal%nscatt_gas = 0.0_c_double; tal%nscatt_dust = 0.0_c_double
    tal%counters = c_lart_counters(0.0_c_double, 0.0_c_double, 0.0_c_double, 0.0_c_double, 0.0_c_double, 0.0_c_double, &
                                   0.0_c_double, 0.0_c_double)
    call check(lart_gpu_fetch(handle, tal))
    par%nscatt_gas  = par%nscatt_gas  + tal%nscatt_gas       ! src/run_simulation_mod.f90:189-191
    par%nscatt_dust = par%nscatt_dust + tal%nscatt_dust
    call check(lart_gpu_destroy(handle))
  end subroutine run_gpu

end module lart_gpu_shim
