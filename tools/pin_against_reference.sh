#!/bin/bash
# tools/pin_against_reference.sh REF [WORK] -- pin the oracle (and, with a B200, the library) against the REFERENCE ITSELF.
#
# The graft image has no Fortran compiler, no Julia and no FFTW, so parity of the UA loop is "unpinned" there.
# On any machine with gfortran + FFTW3 (and optionally julia with the reference's dependencies instantiated) this script
#   1. compiles the reference's own Fortran modules straight from its checkout REF/fortran (nothing is copied; this also
#      sidesteps the reference Makefile, which lists *.f90 for files that are *.F90 and links HDF5 it does not need here),
#   2. links them with tools/pin/pin_driver.F90 (bupdate.F90's call sequence, particles read from a file, energy of
#      src/poisson.jl:80-81 recorded after every solve),
#   3. runs it -- and REF's Julia package through tools/pin/pin_julia.jl if `julia` exists -- on the referee inputs
#      (eps = 1e-1 ... 1e-5) and on BASELINE config 1 as shipped (204 800 particles, ntau 16, 128 x 64, 8 steps),
#   4. compares x, v and the energy history with oracle/ (Fortran wrap for the Fortran run, Julia wrap for the Julia run) and
#      with libuapic_b200 when a GPU is present, at 1e-10 (v: 1e-10 * max(1, 1e-3/eps), see tests/test_gpu_referee.py),
#   5. writes the Fortran outputs to tests/golden/pinned_*.npz -- true reference-held vectors; tests/test_oracle.py and
#      tests/test_gpu_session.py pick up every tests/golden/*.npz automatically.
# Exit code 0 = everything within tolerance.
set -euo pipefail
REF=${1:?usage: pin_against_reference.sh /path/to/UAPIC.jl [workdir]}
WORK=${2:-/tmp/uapic_pin}
HERE=$(cd "$(dirname "$0")" && pwd)
FC=${FC:-gfortran}
FFTW_INC=${FFTW_INC:-$( (pkg-config --variable=includedir fftw3 2>/dev/null) || echo /usr/include)}
FFTW_LIB=${FFTW_LIB:-$( (pkg-config --libs fftw3 2>/dev/null) || echo -lfftw3)}
command -v "$FC" >/dev/null || { echo "no Fortran compiler ($FC): parity stays unpinned on this machine"; exit 3; }
[ -f "$FFTW_INC/fftw3.f03" ] || { echo "fftw3.f03 not found under $FFTW_INC (set FFTW_INC)"; exit 3; }
mkdir -p "$WORK"
# -ffp-contract=off: the oracle restates the Fortran without FMA contraction (gfortran on baseline x86-64 does not fuse either)
FFLAGS="-O2 -ffp-contract=off -cpp -I$FFTW_INC -J$WORK"
for m in meshfields.F90 particles.F90 ua_type.F90 compute_rho_m6.F90 interpolation_m6.F90 poisson_2d.f90 ua_steps.F90; do
    $FC $FFLAGS -c "$REF/fortran/$m" -o "$WORK/${m%.*}.o"
done
$FC $FFLAGS "$HERE/pin/pin_driver.F90" "$WORK"/{meshfields,particles,ua_type,compute_rho_m6,interpolation_m6,poisson_2d,ua_steps}.o $FFTW_LIB -o "$WORK/pin_driver"
python "$HERE/pin/pin_io.py" inputs "$WORK" | while read -r name; do
    "$WORK/pin_driver" "$WORK/$name.in" "$WORK/$name.fortran.out" > "$WORK/$name.fortran.log"
    if command -v julia >/dev/null; then
        julia --project="$REF" "$HERE/pin/pin_julia.jl" "$REF" "$WORK/$name.in" "$WORK/$name.julia.out" > "$WORK/$name.julia.log" 2>&1 || echo "julia run of $name failed (see $WORK/$name.julia.log)"
    fi
done
# 6. the external-field program: the reference's efd.f90 with its debug truncation lifted IN A TEMPORARY COPY (particle loop over
#    all particles, the mid-way `stop` and the per-particle prints dropped) -- it then prints sum(v) + the constants of its
#    line 481, which must vanish, and so must `python -c "import uapic_b200 as u; print(u.efd(16)[2])"` on a B200.
sed -e 's/^do m=1,1!nbpart/do m=1,nbpart/' -e '/^    stop$/d' -e '/^    print\*,/d' "$REF/fortran/efd.f90" > "$WORK/efd_all_particles.f90"
$FC $FFLAGS -c "$REF/fortran/fft.f90" -o "$WORK/fft.o"
$FC $FFLAGS "$WORK/efd_all_particles.f90" "$WORK"/{fft,meshfields,particles,compute_rho_m6}.o $FFTW_LIB -o "$WORK/efd_all" \
    && { echo "reference efd.f90 over all particles: sum(v) + printed constants (expect ~1e-10):"; "$WORK/efd_all" | tail -2; } \
    || echo "efd.f90 did not build here (it needs only fft.f90, meshfields, particles, compute_rho_m6 and FFTW)"
python "$HERE/pin/pin_io.py" compare "$WORK"
