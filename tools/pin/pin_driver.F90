! pin_driver.F90 -- OUR driver around the REFERENCE's own Fortran modules (compiled from the reference checkout where it lies;
! nothing of the reference is copied into this repository).  It is fortran/bupdate.F90's program with two changes:
!   * the particles come from a binary file (the same doubles the oracle and the GPU get) instead of init_particles_2d's
!     RANDOM_NUMBER stream, and nx, ny, ntau, eps, nstep are read from that file;
!   * after every solve_poisson the electric energy sum(e1^2+e2^2)*dx*dy over the ghosted array is recorded -- the definition
!     of src/poisson.jl:80-81 (the Fortran has none) -- and x, v, the energy history and the last E mesh are written out.
! The call sequence is bupdate.F90:89-128 verbatim, third (dead) interpolation included.
! Built and run by tools/pin_against_reference.sh on a machine with gfortran + FFTW3; cannot be built in the graft image.
program pin_driver

    use mesh_fields_m
    use particles_m
    use poisson_2d_m
    use m6_interpolation_m
    use m6_compute_rho_m
    use ua_steps_m

    implicit none

    integer(8) :: npt
    integer    :: nx, ny, nta, nst, it, ie
    real(8)    :: epsv, dtm, xmax, ymax, w
    type(mesh_t)      :: msh
    type(fields_2d_t) :: fld
    type(particles_t) :: pt
    type(poisson_t)   :: psn
    type(ua_t)        :: uat
    complex(8), allocatable :: xtau(:,:,:), xhat(:,:,:), ytau(:,:,:), yhat(:,:,:), fxh(:,:,:), fyh(:,:,:), gxh(:,:,:), gyh(:,:,:)
    real(8), allocatable :: etau(:,:,:), energy(:)
    character(len=512) :: fin, fout

    call get_command_argument(1, fin)
    call get_command_argument(2, fout)
    open(10, file=trim(fin), access='stream', form='unformatted', status='old')
    read(10) npt, nx, ny, nta, nst
    read(10) epsv, dtm, xmax, ymax, w
    pt%nbpart = npt
    allocate(pt%x(2,npt), pt%v(2,npt), pt%e(2,npt), pt%b(npt), pt%t(npt))
    read(10) pt%x
    read(10) pt%v
    close(10)
    pt%w = w

    pi = 4d0 * atan(1d0)
    call init_mesh( msh, 0d0, xmax, nx, 0d0, ymax, ny )
    call init_fields( fld, msh )
    call init_poisson( psn, msh )
    call init_ua( uat, nta, epsv, npt )

    allocate(etau(nta,2,npt), xtau(nta,2,npt), xhat(nta,2,npt), ytau(nta,2,npt), yhat(nta,2,npt))
    allocate(fxh(nta,2,npt), fyh(nta,2,npt), gxh(nta,2,npt), gyh(nta,2,npt))
    allocate(energy(1 + 2*nst))
    ie = 0

    call compute_rho_m6_real( fld, pt )
    call solve_poisson( psn, fld );  call record()
    call interpolate_eb_m6_real( pt, fld )

    do it = 1, nst
        call preparation( uat, dtm, pt, xtau, ytau)
        call interpolation( pt, etau, fld, uat, xtau)
        call compute_f( fxh, fyh, uat, pt, xtau, ytau, etau )
        call ua_step1( xtau, xhat, uat, pt, fxh )
        call ua_step1( ytau, yhat, uat, pt, fyh )
        call deposition( pt, fld, uat, xtau)
        call solve_poisson( psn, fld );  call record()
        call interpolation( pt, etau, fld, uat, xtau)
        call compute_f( gxh, gyh, uat, pt, xtau, ytau, etau )
        call ua_step2( xtau, xhat, uat, pt, fxh, gxh )
        call ua_step2( ytau, yhat, uat, pt, fyh, gyh )
        call deposition( pt, fld, uat, xtau)
        call solve_poisson( psn, fld );  call record()
        call interpolation( pt, etau, fld, uat, xtau)
        call compute_v( uat, pt, ytau, yhat )
    end do

    open(11, file=trim(fout), access='stream', form='unformatted', status='replace')
    write(11) npt, nx, ny, nta, nst
    write(11) pt%x
    write(11) pt%v
    write(11) energy
    write(11) fld%e
    close(11)

contains

    subroutine record()
        ie = ie + 1
        energy(ie) = sum(fld%e(1,:,:)**2 + fld%e(2,:,:)**2) * msh%dx * msh%dy
    end subroutine record

end program pin_driver
