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

    integer(8) :: nbpart
    integer    :: nx, ny, ntau, nstep, istep, ie
    real(8)    :: eps, dt, xmax, ymax, w
    type(mesh_t)      :: mesh
    type(fields_2d_t) :: fields
    type(particles_t) :: particles
    type(poisson_t)   :: poisson
    type(ua_t)        :: ua
    complex(8), allocatable :: xt(:,:,:), xf(:,:,:), yt(:,:,:), yf(:,:,:), fx(:,:,:), fy(:,:,:), gx(:,:,:), gy(:,:,:)
    real(8), allocatable :: et(:,:,:), energy(:)
    character(len=512) :: fin, fout

    call get_command_argument(1, fin)
    call get_command_argument(2, fout)
    open(10, file=trim(fin), access='stream', form='unformatted', status='old')
    read(10) nbpart, nx, ny, ntau, nstep
    read(10) eps, dt, xmax, ymax, w
    particles%nbpart = nbpart
    allocate(particles%x(2,nbpart), particles%v(2,nbpart), particles%e(2,nbpart), particles%b(nbpart), particles%t(nbpart))
    read(10) particles%x
    read(10) particles%v
    close(10)
    particles%w = w

    pi = 4d0 * atan(1d0)
    call init_mesh( mesh, 0d0, xmax, nx, 0d0, ymax, ny )
    call init_fields( fields, mesh )
    call init_poisson( poisson, mesh )
    call init_ua( ua, ntau, eps, nbpart )

    allocate(et(ntau,2,nbpart), xt(ntau,2,nbpart), xf(ntau,2,nbpart), yt(ntau,2,nbpart), yf(ntau,2,nbpart))
    allocate(fx(ntau,2,nbpart), fy(ntau,2,nbpart), gx(ntau,2,nbpart), gy(ntau,2,nbpart))
    allocate(energy(1 + 2*nstep))
    ie = 0

    call compute_rho_m6_real( fields, particles )
    call solve_poisson( poisson, fields );  call record()
    call interpolate_eb_m6_real( particles, fields )

    do istep = 1, nstep
        call preparation( ua, dt, particles, xt, yt)
        call interpolation( particles, et, fields, ua, xt)
        call compute_f( fx, fy, ua, particles, xt, yt, et )
        call ua_step1( xt, xf, ua, particles, fx )
        call ua_step1( yt, yf, ua, particles, fy )
        call deposition( particles, fields, ua, xt)
        call solve_poisson( poisson, fields );  call record()
        call interpolation( particles, et, fields, ua, xt)
        call compute_f( gx, gy, ua, particles, xt, yt, et )
        call ua_step2( xt, xf, ua, particles, fx, gx )
        call ua_step2( yt, yf, ua, particles, fy, gy )
        call deposition( particles, fields, ua, xt)
        call solve_poisson( poisson, fields );  call record()
        call interpolation( particles, et, fields, ua, xt)
        call compute_v( ua, particles, yt, yf )
    end do

    open(11, file=trim(fout), access='stream', form='unformatted', status='replace')
    write(11) nbpart, nx, ny, ntau, nstep
    write(11) particles%x
    write(11) particles%v
    write(11) energy
    write(11) fields%e
    close(11)

contains

    subroutine record()
        ie = ie + 1
        energy(ie) = sum(fields%e(1,:,:)**2 + fields%e(2,:,:)**2) * mesh%dx * mesh%dy
    end subroutine record

end program pin_driver
