! movfem_cuda.f90 -- ISO_C_BINDING shim: calls libmovfem_b200.so from MoVFEM_3DMT.f90.
!
! NOT compiled in this repository's environment (no Fortran compiler in the image); it is the binding a
! maintainer adds to MoVFEM_3DMT/src and lists in the Makefile.  Every interface below is the Fortran view
! of one prototype in include/movfem_b200.h.  See INTEGRATION.md for the three edits to MoVFEM_3DMT.f90.
!
! It replaces, on rank 0 only (MoVFEM_3DMT.f90:63):
!   ga_init -> ga_cgne / ga_nzindx                      (global_assembly.f90:38-39)   by movfem_cuda_init
!   global_vfem + find_zeros / rem_zeros                (MoVFEM_3DMT.f90:82-97)       by movfem_cuda_assemble
module movfem_cuda
    use, intrinsic :: iso_c_binding
    use kind_param
    implicit none
    private
    public :: movfem_cuda_init, movfem_cuda_assemble, movfem_cuda_finalize, movfem_nz_upper, movfem_cuda_set_boundary_model

    ! struct movfem_desc of include/movfem_b200.h, field for field
    type, bind(C) :: movfem_desc
        integer(c_int32_t) :: g_nx, g_ny, g_nz, nord, mn, me, nextd, nzl_top
        integer(c_int32_t) :: dirichlet, bd_inimod, gpml_sch, sym, ndir, pe_sch
        real(c_double)     :: a0, b0, nn
        type(c_ptr)        :: g_xp, g_yp, g_zp, g_mu
        integer(c_int32_t) :: ie_lo, ie_hi
        real(c_double)     :: g_ztop, bd_hsigma
        integer(c_int32_t) :: bd_nl, bd_pad
        real(c_double)     :: bd_lsigma(16), bd_ldz(16)
    end type movfem_desc

    ! Dirichlet boundary models 2/3: the arguments of bd_setmodel (MoVFEM_3DMT.f90:388) are locals of the input reader
    ! and boundary_conds keeps them private, so the driver hands a copy to movfem_cuda_set_boundary_model right there.
    real(c_double), save     :: bdm_hsigma = 0.d0, bdm_lsigma(16) = 0.d0, bdm_ldz(16) = 0.d0
    integer(c_int32_t), save :: bdm_nl = 1

    type(c_ptr), save          :: handle = c_null_ptr
    integer(c_int64_t), save   :: movfem_nz_upper = 0     ! capacity irn/jcn/a need (<= nnze)
    integer(c_int32_t), parameter :: MODE_T2 = 0, MODE_KEEP_PATTERN = 256      ! MOVFEM_MODE_T2, MOVFEM_MODE_KEEP_PATTERN

    interface
        integer(c_int) function movfem_create(desc, device, h) bind(C, name='movfem_create')
            import :: c_int, c_ptr, movfem_desc
            type(movfem_desc), intent(in) :: desc
            integer(c_int), value         :: device
            type(c_ptr), intent(out)      :: h
        end function
        subroutine movfem_destroy(h) bind(C, name='movfem_destroy')
            import :: c_ptr
            type(c_ptr), value :: h
        end subroutine
        integer(c_int) function movfem_sizes(h, nne, nnze_full, nz_upper) bind(C, name='movfem_sizes')
            import :: c_int, c_ptr, c_int32_t, c_int64_t
            type(c_ptr), value              :: h
            integer(c_int32_t), intent(out) :: nne
            integer(c_int64_t), intent(out) :: nnze_full, nz_upper
        end function
        integer(c_int) function movfem_get_gne(h, gne) bind(C, name='movfem_get_gne')
            import :: c_int, c_ptr, c_int32_t
            type(c_ptr), value                :: h
            integer(c_int32_t), intent(out)   :: gne(*)          ! gne(ne,me), column-major as in Fortran
        end function
        integer(c_int) function movfem_assemble(h, freq_index, omega, g_sigma, irn, jcn, a, rhs, nz_out, mode) &
                bind(C, name='movfem_assemble')
            import :: c_int, c_ptr, c_int32_t, c_int64_t, c_double, c_double_complex
            type(c_ptr), value                         :: h
            integer(c_int32_t), value                  :: freq_index
            real(c_double), value                      :: omega
            complex(c_double_complex), intent(in)      :: g_sigma(*)   ! (6,g_npt)
            integer(c_int32_t), intent(out)            :: irn(*), jcn(*)
            complex(c_double_complex), intent(out)     :: a(*), rhs(*)
            integer(c_int64_t), intent(out)            :: nz_out
            integer(c_int32_t), value                  :: mode
        end function
        function movfem_last_error(h) bind(C, name='movfem_last_error') result(msg)
            import :: c_ptr
            type(c_ptr), value :: h
            type(c_ptr)        :: msg
        end function
    end interface

contains

    ! call movfem_cuda_set_boundary_model(h_sigma, nl, l_sigma, l_dz)   next to   call bd_setmodel(h_sigma,nl,l_sigma,l_dz)
    subroutine movfem_cuda_set_boundary_model(h_sigma, nl, l_sigma, l_dz)
        real(kind=double), intent(in) :: h_sigma, l_sigma(nl), l_dz(nl-1)
        integer, intent(in)           :: nl
        if (nl > 16) call die('more than 16 boundary layers', -1)
        bdm_hsigma = h_sigma; bdm_nl = nl
        bdm_lsigma(1:nl) = l_sigma(1:nl)
        if (nl > 1) bdm_ldz(1:nl-1) = l_dz(1:nl-1)
    end subroutine movfem_cuda_set_boundary_model

    ! Called from ga_init (global_assembly.f90:26) in place of ga_cgne / ga_nzindx:
    !     call movfem_cuda_init(sym, nne, nnze, gne)
    ! It fills the module variables of global_assembly the rest of the program consumes -- nne, nnze (MoVFEM_3DMT.f90:72-78),
    ! gne (solution.f90:331-336) -- which arrive as ARGUMENTS: this module must not `use global_assembly`, because
    ! global_assembly uses this module (a circular module dependency does not compile).  Modules used here: geometry, n_fem,
    ! v_fem, problem, boundary_conds, kind_param -- none of them uses global_assembly (tests/test_c_harness.py sorts the graph).
    subroutine movfem_cuda_init(sym, nne, nnze, gne)
        use geometry, only: g_nx, g_ny, g_nz, g_nordx, nextd, g_nsf, g_nzl, g_xp, g_yp, g_zp, g_mu, g_ztop
        use n_fem, only: nf_mn
        use v_fem, only: vf_me
        use problem, only: ndir, pe_sch
        use boundary_conds, only: dirichlet, bd_inimod, gpml_sch, a0, b0, nn
        logical, intent(in)               :: sym
        integer, intent(out)              :: nne, nnze
        integer, allocatable, intent(out) :: gne(:,:)
        type(movfem_desc) :: d
        integer(c_int) :: rc
        integer(c_int32_t) :: nne_c
        integer(c_int64_t) :: nnze_c
        d%g_nx = g_nx; d%g_ny = g_ny; d%g_nz = g_nz; d%nord = g_nordx
        d%mn = nf_mn; d%me = vf_me; d%nextd = nextd; d%nzl_top = g_nzl(g_nsf)
        d%dirichlet = merge(1, 0, dirichlet); d%bd_inimod = bd_inimod; d%gpml_sch = gpml_sch
        d%sym = merge(1, 0, sym); d%ndir = ndir; d%pe_sch = pe_sch
        d%a0 = a0; d%b0 = b0; d%nn = nn
        d%g_xp = c_loc(g_xp); d%g_yp = c_loc(g_yp); d%g_zp = c_loc(g_zp); d%g_mu = c_loc(g_mu)   ! arrays need TARGET
        d%ie_lo = 0; d%ie_hi = 0
        d%g_ztop = g_ztop; d%bd_hsigma = bdm_hsigma; d%bd_nl = bdm_nl; d%bd_pad = 0
        d%bd_lsigma = bdm_lsigma; d%bd_ldz = bdm_ldz
        rc = movfem_create(d, 0_c_int, handle)
        if (rc /= 0) call die('movfem_create', rc)
        rc = movfem_sizes(handle, nne_c, nnze_c, movfem_nz_upper)
        if (rc /= 0) call die('movfem_sizes', rc)
        if (nnze_c > huge(nnze)) call die('nnze exceeds default integer (SURVEY Q16): size arrays with movfem_nz_upper', -6)
        nne = nne_c; nnze = int(nnze_c)
        allocate(gne((g_nx-1)*(g_ny-1)*(g_nz-1), vf_me))
        rc = movfem_get_gne(handle, gne)
        if (rc /= 0) call die('movfem_get_gne', rc)
    end subroutine movfem_cuda_init

    ! Replaces  call global_vfem(irn,jcn,a,rhs)  and the find_zeros/rem_zeros block (MoVFEM_3DMT.f90:82-97).
    ! ii is the loop index of the frequency loop (MoVFEM_3DMT.f90:62); nz returns mumps_par%nz.
    ! keep_pattern (optional, .true. when irn/jcn are the SAME arrays the previous call filled, i.e. their allocation was hoisted
    ! out of the frequency loop): the static pattern is not sent again, a third of the device-to-host traffic.
    subroutine movfem_cuda_assemble(ii, irn, jcn, a, rhs, nz, keep_pattern)
        use geometry, only: omega, g_sigma
        integer, intent(in)                          :: ii
        integer, intent(inout)                       :: irn(*), jcn(*)
        complex(kind=double), intent(inout)          :: a(*), rhs(*)
        integer, intent(out)                         :: nz
        logical, intent(in), optional                :: keep_pattern
        integer(c_int64_t) :: nz_c
        integer(c_int) :: rc
        integer(c_int32_t) :: mode
        mode = MODE_T2
        if (present(keep_pattern)) then
            if (keep_pattern) mode = MODE_T2 + MODE_KEEP_PATTERN
        end if
        rc = movfem_assemble(handle, int(ii, c_int32_t), omega, g_sigma, irn, jcn, a, rhs, nz_c, mode)
        if (rc /= 0) call die('movfem_assemble', rc)
        nz = int(nz_c)
    end subroutine movfem_cuda_assemble

    subroutine movfem_cuda_finalize()
        if (c_associated(handle)) call movfem_destroy(handle)
        handle = c_null_ptr
    end subroutine movfem_cuda_finalize

    ! the reference stops on inconsistencies (global_assembly.f90:56-57, n_fem.f90:374-377, problem.f90:260-271)
    subroutine die(what, rc)
        character(*), intent(in) :: what
        integer(c_int), intent(in) :: rc
        character(kind=c_char), pointer :: s(:)
        type(c_ptr) :: p
        integer :: i
        print *, trim(what), ' failed with code ', rc
        if (c_associated(handle)) then
            p = movfem_last_error(handle)
            if (c_associated(p)) then
                call c_f_pointer(p, s, [512])
                do i = 1, 512
                    if (s(i) == c_null_char) exit
                    write(*, '(a)', advance='no') s(i)
                end do
                print *
            end if
        end if
        stop
    end subroutine die
end module movfem_cuda

! Geomodel -> grid nodes (SURVEY 8f rank 4): drop-in for `call innermodel_gqg(mx,my,mz,xm,ym,zm,isigma,imu,ijsigma,ijmu,
! sigma,mu)` at geometry.f90:99.  A separate module that does NOT use geometry (geometry.f90 itself calls it), so the
! mesh arrives as arguments: the caller allocates g_sigma(6,g_npt), g_mu(6,g_npt) exactly as innermodel_gqg does
! (geometry.f90:824) and passes its module variables.
!
!   ! geometry.f90:99, was: call innermodel_gqg(mx,my,mz,xm,ym,zm,isigma, imu, ijsigma,ijmu, sigma,mu)
!   allocate(g_sigma(6,g_npt), g_mu(6,g_npt))
!   call movfem_cuda_innermodel(g_nx,g_ny,g_nz,g_nordx,nextd,g_nzl(g_nsf),g_nzl(g_nsf-1),g_xp,g_yp,g_zp,omega, &
!                               mx,my,mz,xm,ym,zm,isigma,imu,ijsigma,ijmu,sigma,mu,g_sigma,g_mu)
module movfem_cuda_geo
    use, intrinsic :: iso_c_binding
    use kind_param
    implicit none
    private
    public :: movfem_cuda_innermodel

    type, bind(C) :: movfem_desc_geo        ! struct movfem_desc of include/movfem_b200.h, field for field
        integer(c_int32_t) :: g_nx, g_ny, g_nz, nord, mn, me, nextd, nzl_top
        integer(c_int32_t) :: dirichlet, bd_inimod, gpml_sch, sym, ndir, pe_sch
        real(c_double)     :: a0, b0, nn
        type(c_ptr)        :: g_xp, g_yp, g_zp, g_mu
        integer(c_int32_t) :: ie_lo, ie_hi
        real(c_double)     :: g_ztop, bd_hsigma
        integer(c_int32_t) :: bd_nl, bd_pad
        real(c_double)     :: bd_lsigma(16), bd_ldz(16)
    end type movfem_desc_geo

    type, bind(C) :: movfem_geomodel        ! struct movfem_geomodel
        integer(c_int32_t) :: mx, my, mz, isigma, imu, nzl_air
        integer(c_int32_t) :: ijsigma(2,9), ijmu(2,9)      ! C [9][2]: (row, col) of component i in column i
        type(c_ptr)        :: xm, ym, zm, sigma, mu
    end type movfem_geomodel

    interface
        integer(c_int) function movfem_geo_innermodel(mesh, device, gm, omega, g_sigma, g_mu, ms_device) &
                bind(C, name='movfem_geo_innermodel')
            import :: c_int, c_int32_t, c_double, c_double_complex, c_ptr, movfem_desc_geo, movfem_geomodel
            type(movfem_desc_geo), intent(in) :: mesh
            integer(c_int32_t), value :: device
            type(movfem_geomodel), intent(in) :: gm
            real(c_double), value :: omega
            complex(c_double_complex), intent(out) :: g_sigma(6,*)
            real(c_double), intent(out) :: g_mu(6,*)
            type(c_ptr), value :: ms_device
        end function
    end interface

    contains

    subroutine movfem_cuda_innermodel(g_nx,g_ny,g_nz,nord,nextd,nzl_top,nzl_air,g_xp,g_yp,g_zp,omega, &
                                      mx,my,mz,xm,ym,zm,isigma,imu,ijsigma,ijmu,sigma,mu,g_sigma,g_mu)
        integer, intent(in) :: g_nx,g_ny,g_nz,nord,nextd,nzl_top,nzl_air,mx,my,mz,isigma,imu
        integer, intent(in) :: ijsigma(isigma,2), ijmu(imu,2)
        real(kind=double), intent(in), target :: g_xp(*), g_yp(*), g_zp(*), xm(mx), ym(my), zm(*), sigma(isigma,*), mu(imu,*)
        real(kind=double), intent(in) :: omega
        complex(kind=double), intent(out) :: g_sigma(6,*)
        real(kind=double), intent(out) :: g_mu(6,*)
        type(movfem_desc_geo) :: d
        type(movfem_geomodel) :: gm
        integer :: i
        integer(c_int) :: rc
        d%g_nx = g_nx; d%g_ny = g_ny; d%g_nz = g_nz; d%nord = nord; d%nextd = nextd; d%nzl_top = nzl_top
        d%mn = 0; d%me = 0; d%dirichlet = 0; d%bd_inimod = 1; d%gpml_sch = 0; d%sym = 1; d%ndir = 2; d%pe_sch = 1
        d%a0 = 0.d0; d%b0 = 0.d0; d%nn = 0.d0; d%ie_lo = 0; d%ie_hi = 0; d%g_ztop = 0.d0; d%bd_hsigma = 0.d0; d%bd_nl = 1; d%bd_pad = 0
        d%bd_lsigma = 0.d0; d%bd_ldz = 0.d0
        d%g_xp = c_loc(g_xp); d%g_yp = c_loc(g_yp); d%g_zp = c_loc(g_zp); d%g_mu = c_null_ptr
        gm%mx = mx; gm%my = my; gm%mz = mz; gm%isigma = isigma; gm%imu = imu; gm%nzl_air = nzl_air
        gm%ijsigma = 0; gm%ijmu = 0
        do i = 1, isigma
            gm%ijsigma(1,i) = ijsigma(i,1); gm%ijsigma(2,i) = ijsigma(i,2)
        end do
        do i = 1, imu
            gm%ijmu(1,i) = ijmu(i,1); gm%ijmu(2,i) = ijmu(i,2)
        end do
        gm%xm = c_loc(xm); gm%ym = c_loc(ym); gm%zm = c_loc(zm); gm%sigma = c_loc(sigma); gm%mu = c_loc(mu)
        rc = movfem_geo_innermodel(d, 0_c_int32_t, gm, omega, g_sigma, g_mu, c_null_ptr)
        if (rc /= 0) then
            print *, 'movfem_geo_innermodel failed with code ', rc
            stop
        end if
    end subroutine movfem_cuda_innermodel
end module movfem_cuda_geo
