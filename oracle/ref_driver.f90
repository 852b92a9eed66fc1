!> ref_driver.f90 -- OpenMP ensemble driver around the UNMODIFIED reference (princemahajan/FLINT).
!!
!! Test infrastructure (see oracle/flint_oracle.cpp): this program is the "real reference" arm of
!! bench.py (`--impl reference`, cpu_baseline.kind = "reference") and the cross-language check of the
!! C++ oracle.  It is this repository's own code; it only USEs the reference's public module FLINT
!! (src/FLINT.f90:84-92) exactly as the reference's tests do (tests/test_CR3BP.f90:339-341,390-404) and is
!! linked against the reference's sources compiled where they lie (`make -C oracle ref`, which needs a
!! Fortran compiler: none exists in the build image, so oracle/_ref stays empty there and bench.py falls
!! back to the C++ port and says so).
!!
!!   flint_ref_driver N START THREADS [DUMPFILE]
!!
!! integrates trajectories START .. START+N-1 of the seeded north-star ensemble (SURVEY.md 8d config 3a:
!! planar CR3BP, y0(1) = 0.994*(1 + 1e-6*u_i), DOP853, rtol 1e-12, atol 1e-15, test_CR3BP event set) with one
!! ERK_class and one DiffEqSys object per thread (both carry mutable state), `schedule(dynamic,64)`, and
!! prints "N seconds threads".  DUMPFILE (optional, stream access): N records of
!! (status, nSteps, nAccept, nReject, nFCalls as int32; Xf, Yf(1:6), event x, event y(1:6) as real64) for
!! tests/test_reference_arm.py to compare with the oracle bit for bit.
module RefCR3BP
    use FLINT, WP => FLINT_WP
    use, intrinsic :: IEEE_ARITHMETIC
    implicit none

    real(WP), parameter :: mu0 = 0.012277471_WP                      ! tests/test_CR3BP.f90:33

    type, extends(DiffEqSys) :: RefSys
        real(WP) :: mu = mu0
    contains
        procedure :: F => RefDE
    end type RefSys

contains

    !> The planar CR3BP right-hand side in the reference test's own expression order (tests/test_CR3BP.f90:167-186)
    function RefDE(me, X, Y, Params)
        class(RefSys), intent(inout) :: me
        real(WP), intent(in) :: X
        real(WP), intent(in), dimension(me%n) :: Y
        real(WP), intent(in), dimension(:), optional :: Params
        real(WP), dimension(size(Y)) :: RefDE
        real(WP) :: r1cube, r2cube
        r1cube = norm2([(Y(1) + me%mu), Y(2)])**3
        r2cube = norm2([(Y(1) - (1.0_WP - me%mu)), Y(2)])**3
        RefDE(1:3) = Y(4:6)
        RefDE(4) = Y(1) + 2.0_WP*Y(5) - (1.0_WP - me%mu)*(Y(1) + me%mu)/r1cube - me%mu*(Y(1) - (1.0_WP - me%mu))/r2cube
        RefDE(5) = Y(2) - 2.0_WP*Y(4) - (1.0_WP - me%mu)*Y(2)/r1cube - me%mu*Y(2)/r2cube
        RefDE(6) = 0.0_WP
    end function RefDE

    !> The three sample events of tests/test_CR3BP.f90:97-164
    subroutine RefEvents(me, X, Y, EvalEvents, Value, Direction, LocEvent, LocEventAction)
        class(DiffEqSys), intent(inout) :: me
        real(WP), intent(in) :: X
        real(WP), dimension(:), intent(inout) :: Y
        integer, dimension(:), intent(in) :: EvalEvents
        real(WP), dimension(:), intent(out) :: Value
        integer, dimension(:), intent(out) :: Direction
        integer, intent(in), optional :: LocEvent
        integer(kind(FLINT_EVENTACTION_CONTINUE)), intent(out), optional :: LocEventAction
        if (EvalEvents(1) == 1) then
            if (X > 0.5_WP .and. X < 0.5005_WP) then
                Value(1) = 1
            else
                Value(1) = -1
            end if
            Direction(1) = 0
        end if
        if (EvalEvents(2) == 1) then
            if (Y(1) < 0.0_WP) then
                Value(2) = IEEE_VALUE(1.0, IEEE_QUIET_NAN)
            else
                Value(2) = Y(2)
            end if
            Direction(2) = 1
        end if
        if (EvalEvents(3) == 1) then
            if (Y(1) > 0.0_WP) then
                Value(3) = IEEE_VALUE(1.0, IEEE_QUIET_NAN)
            else
                Value(3) = Y(2)
            end if
            Direction(3) = 1
        end if
        if (present(LocEvent) .and. present(LocEventAction)) then
            if (LocEvent == 1 .or. LocEvent == 2) then
                LocEventAction = ior(FLINT_EVENTACTION_CONTINUE, FLINT_EVENTACTION_MASK)
            else if (LocEvent == 3) then
                Y = [0.994_WP, 0.0_WP, 0.0_WP, 0.0_WP, -2.00158510637908252240537862224_WP, 0.0_WP]
                LocEventAction = ior(FLINT_EVENTACTION_CHANGESOLY, FLINT_EVENTACTION_MASK)
            end if
        end if
    end subroutine RefEvents

    !> u in [0,1): splitmix64 of seed xor (index*0x9E3779B97F4A7C15 + comp), top 53 bits -- the generator of
    !! flint_b200/problems.py:uniform01 (two's-complement wrap-around arithmetic: compile with -fwrapv)
    function uniform01(idx, comp) result(u)
        use, intrinsic :: iso_fortran_env, only: int64
        integer(int64), intent(in) :: idx
        integer, intent(in) :: comp
        real(WP) :: u
        integer(int64), parameter :: GOLD = -7046029254386353131_int64          ! 0x9E3779B97F4A7C15
        integer(int64), parameter :: M1 = -4658895280553007687_int64            ! 0xBF58476D1CE4E5B9
        integer(int64), parameter :: M2 = -7723592293110705685_int64            ! 0x94D049BB133111EB
        integer(int64), parameter :: SEED = 20240607_int64
        integer(int64) :: z
        z = ieor(SEED, idx*GOLD + int(comp, int64))
        z = z + GOLD
        z = ieor(z, shiftr(z, 30))*M1
        z = ieor(z, shiftr(z, 27))*M2
        z = ieor(z, shiftr(z, 31))
        u = real(shiftr(z, 11), WP)*(1.0_WP/9007199254740992.0_WP)
    end function uniform01

end module RefCR3BP


program flint_ref_driver
    use RefCR3BP
    use omp_lib
    use, intrinsic :: iso_fortran_env, only: int32, int64
    implicit none

    integer(int64) :: N, START, i
    integer :: nthreads, nargs, dump_unit
    character(len=256) :: arg, dumpfile
    logical :: dump
    real(WP), parameter :: rtol = 1.0e-12_WP, atol = rtol*1.0e-3_WP                    ! tests/test_CR3BP.f90:248
    real(WP), parameter :: period = 17.0652165601579625588917206249_WP                 ! :336
    real(WP), parameter :: y0nom(6) = [0.994_WP, 0.0_WP, 0.0_WP, 0.0_WP, -2.00158510637908252240537862224_WP, 0.0_WP]
    integer, parameter :: MAX_STEPS = 100000
    integer(int32), allocatable :: ires(:, :)
    real(WP), allocatable :: dres(:, :)
    real(WP) :: t0, t1

    nargs = command_argument_count()
    if (nargs < 3) then
        write(*, *) 'usage: flint_ref_driver N START THREADS [DUMPFILE]'
        stop 2
    end if
    call get_command_argument(1, arg); read(arg, *) N
    call get_command_argument(2, arg); read(arg, *) START
    call get_command_argument(3, arg); read(arg, *) nthreads
    dump = nargs >= 4
    if (dump) call get_command_argument(4, dumpfile)
    if (nthreads > 0) call omp_set_num_threads(nthreads)
    allocate(ires(5, N), dres(15, N))
    ires = 0
    dres = 0.0_WP

    t0 = omp_get_wtime()
    !$omp parallel default(shared)
    block
        type(ERK_class) :: erk                      ! one integrator object and one system per thread: both carry mutable state
        type(RefSys) :: sys
        real(WP) :: y0(6), yf(6), xf, stepsz
        real(WP), allocatable :: EventStates(:, :)
        integer :: stiff, st, nsteps, nacc, nrej, nfc
        integer(int64) :: k
        sys%n = 6
        sys%m = 3
        sys%G => RefEvents
        call erk%Init(sys, MAX_STEPS, Method=ERK_DOP853, ATol=[atol], RTol=[rtol], InterpOn=.false., EventsOn=.true., &
                      EventStepSz=[real(WP) :: 1.0e-5_WP, 0.0_WP, 0.0_WP], EventTol=[1.0e-3_WP, 1.0e-9_WP, 1.0e-9_WP])
        !$omp do schedule(dynamic, 64)
        do k = 1, N
            y0 = y0nom
            y0(1) = y0nom(1)*(1.0_WP + 1.0e-6_WP*uniform01(START + k - 1, 0))
            xf = 2.5_WP*period
            stepsz = 0.0_WP
            stiff = 1
            call erk%Integrate(0.0_WP, y0, xf, yf, StepSz=stepsz, UseConstStepSz=.false., IntStepsOn=.false., &
                               EventStates=EventStates, EventMask=[.false., .true., .false.], StiffTest=stiff)
            call erk%Info(st, nSteps=nsteps, nAccept=nacc, nReject=nrej, nFCalls=nfc)
            ires(:, k) = [int(st, int32), int(nsteps, int32), int(nacc, int32), int(nrej, int32), int(nfc, int32)]
            dres(1, k) = xf
            dres(2:7, k) = yf
            if (allocated(EventStates)) then
                if (size(EventStates, 2) >= 1) dres(8:14, k) = EventStates(1:7, 1)
                dres(15, k) = real(size(EventStates, 2), WP)
            end if
        end do
        !$omp end do
    end block
    !$omp end parallel
    t1 = omp_get_wtime()

    write(*, '(I0,1X,ES16.8,1X,I0)') N, t1 - t0, omp_get_max_threads()
    if (dump) then
        open(newunit=dump_unit, file=trim(dumpfile), access='stream', form='unformatted', status='replace')
        do i = 1, N
            write(dump_unit) ires(:, i), dres(:, i)
        end do
        close(dump_unit)
    end if
end program flint_ref_driver
