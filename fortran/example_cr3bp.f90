!> example_cr3bp.f90 -- the reference's tests/test_CR3BP.f90 (Init :390-392, Integrate :399-404, Info) with ONE
!! changed line: `use FLINT_GPU` instead of `use FLINT`.  The user's F and G are not passed: the planar CR3BP
!! right-hand side and the three sample events of that test are compiled into the library and selected by id.
!!
!!   gfortran -O2 fortran/flint_gpu.f90 fortran/example_cr3bp.f90 -L flint_b200 -lflint_b200 -Wl,-rpath,$PWD/flint_b200
!!
!! (not compiled in the build container: no Fortran compiler; examples/cr3bp_abi.c makes the same calls from C and
!! IS compiled and run by tests/test_host_abi.py)
program example_cr3bp
    use, intrinsic :: iso_c_binding
    use FLINT_GPU
    implicit none
    integer, parameter :: WP = c_double
    real(WP), parameter :: mu = 0.012277471_WP, Period = 17.0652165601579625588917206249_WP
    type(DiffEqSys_GPU) :: CR3BPDE
    type(ERK_class) :: erk
    real(WP) :: y0(6), yf(6), x0, xf, stepsz, rtol, atol
    real(WP), allocatable :: EventStates(:,:)
    integer :: stiffstatus, status, nSteps, nAccept, nReject, nFCalls
    character(len=:), allocatable :: msg

    y0 = [0.994_WP, 0.0_WP, 0.0_WP, 0.0_WP, -2.00158510637908252240537862224_WP, 0.0_WP]
    x0 = 0.0_WP;  xf = 2.5_WP * Period;  stepsz = 0.0_WP
    rtol = 1.0e-11_WP;  atol = rtol * 1.0e-3_WP

    call CR3BPDE%Create(FLINT_SYS_CR3BP_PLANAR, FLINT_EVSET_CR3BP_SAMPLE3, [mu])     ! n = 6, m = 3
    call erk%Init(CR3BPDE, 100000, Method=ERK_DOP853, ATol=[atol], RTol=[rtol], InterpOn=.FALSE., EventsOn=.TRUE., &
                  EventStepSz=[1.0e-5_WP, 0.0_WP, 0.0_WP], EventTol=[1.0e-3_WP, 1.0e-9_WP, 1.0e-9_WP])
    if (erk%status /= FLINT_SUCCESS) stop 'Init failed'

    stiffstatus = 1
    call erk%Integrate(x0, y0, xf, yf, StepSz=stepsz, UseConstStepSz=.FALSE., EventMask=[.FALSE., .TRUE., .FALSE.], &
                       EventStates=EventStates, StiffTest=stiffstatus)
    call erk%Info(status, msg, nSteps=nSteps, nAccept=nAccept, nReject=nReject, nFCalls=nFCalls)
    print '(A,I0,1X,A)', 'status ', status, msg
    print '(A,I0,A,I0,A,I0,A,I0)', 'steps ', nSteps, ' accepted ', nAccept, ' rejected ', nReject, ' fcalls ', nFCalls
    if (allocated(EventStates)) then
        if (size(EventStates, 2) > 0) print '(A,F12.9,A,F12.9,A,I0)', 'event1 x ', EventStates(1,1), ' y1 ', EventStates(2,1), &
                                                                      ' id ', int(EventStates(8,1))
    end if
end program example_cr3bp
