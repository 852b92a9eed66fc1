!> flint_gpu.f90 -- ISO_C_BINDING shim: the reference's object API on top of libflint_b200.so.
!!
!! A Fortran host program that today says
!!     use FLINT
!!     type(ERK_class) :: erk
!!     call erk%Init(sys, MaxSteps, Method=ERK_DOP853, ATol=[atol], RTol=[rtol], ...)
!!     call erk%Integrate(X0, Y0, Xf, Yf, StepSz, ...)
!! changes the `use` line to `use FLINT_GPU`, declares its system with
!!     type(DiffEqSys_GPU) :: sys ;  call sys%Create(FLINT_SYS_CR3BP_PLANAR, FLINT_EVSET_CR3BP_SAMPLE3, [mu])
!! (F and G are device functors selected by id instead of type-bound procedures, include/flint_b200.h)
!! and gains `IntegrateBatch` for ensembles.  Argument names, optional-ness, defaults and status codes are
!! the reference's (src/FLINT_base.f90:318-533).
!!
!! NOTE: no Fortran compiler exists in the build image, so this file has NOT been compiled here; the same
!! ABI is exercised from C/ctypes by tests/ (flint_b200/api.py is the line-for-line Python twin of this
!! module).  Build on a machine with gfortran:
!!     gfortran -c fortran/flint_gpu.f90 && gfortran host.f90 flint_gpu.o -L flint_b200 -lflint_b200
module FLINT_GPU
    use, intrinsic :: iso_c_binding
    implicit none
    private

    integer, parameter, public :: WP = c_double

    ! ---- enums: values are ABI and equal the reference's (src/FLINT_base.f90:43-61,86-120; src/ERK.f90:56-66)
    integer(c_int), parameter, public :: FLINT_SUCCESS = 0, FLINT_EVENT_TERM = 1, FLINT_STIFF_PROBLEM = 2, FLINT_ERROR = 3, &
        FLINT_ERROR_PARAMS = 4, FLINT_ERROR_MAXSTEPS = 5, FLINT_ERROR_MEMALLOC = 6, FLINT_ERROR_MEMDEALLOC = 7, &
        FLINT_ERROR_DIMENSION = 8, FLINT_ERROR_STEPSZ_TOOSMALL = 9, FLINT_ERROR_INTPSTATES = 10, &
        FLINT_ERROR_INTP_ARRAY = 11, FLINT_ERROR_INTP_OFF = 12, FLINT_ERROR_INIT_REQD = 13, &
        FLINT_ERROR_EVENTPARAMS = 14, FLINT_ERROR_EVENTROOT = 15, FLINT_ERROR_CONSTSTEPSZ = 16
    integer(c_int), parameter, public :: FLINT_EVENTACTION_CONTINUE = 0, FLINT_EVENTACTION_TERMINATE = 1, &
        FLINT_EVENTACTION_CHANGESOLY = 2, FLINT_EVENTACTION_RESTARTINT = 4, FLINT_EVENTACTION_MASK = 128
    integer(c_int), parameter, public :: FLINT_EVENTOPTION_ROOTFINDING = 0, FLINT_EVENTOPTION_STEPBEGIN = 1, &
        FLINT_EVENTOPTION_STEPEND = 2
    integer(c_int), parameter, public :: ERK_DOP853 = 17, ERK_DOP54 = 18, ERK_VERNER98R = 19, ERK_VERNER65E = 20
    integer(c_int), parameter, public :: FLINT_SYS_CR3BP_PLANAR = 1, FLINT_SYS_CR3BP_SPATIAL = 2, FLINT_SYS_LORENZ = 3, &
        FLINT_SYS_TWOBODY = 4
    integer(c_int), parameter, public :: FLINT_EVSET_NONE = 0, FLINT_EVSET_CR3BP_SAMPLE3 = 1, FLINT_EVSET_XY_CROSSING = 2, &
        FLINT_EVSET_APSIS = 3
    integer(c_int), parameter, public :: FLINT_ARITH_STRICT = 0, FLINT_ARITH_FMA = 1
    integer(c_int), parameter, public :: FLINT_MEM_HOST = 0, FLINT_MEM_DEVICE = 1

    ! present-bits of flint_init_opts / flint_int_opts
    integer(c_int), parameter :: I_METHOD = 1, I_INTERP_ON = 2, I_INTERP_STATES = 4, I_MIN_STEP = 8, I_MAX_STEP = 16, &
        I_STEPSZ_PARAMS = 32, I_EVENTS_ON = 64, I_EVENT_STEPSZ = 128, I_EVENT_OPTIONS = 256, I_EVENT_TOL = 512, I_ARITH = 1024
    integer(c_int), parameter :: N_CONST = 1, N_ADVANCE = 2, N_INTSTEPS = 4, N_XINT = 8, N_EVMASK = 16, N_EVSTATES = 32, &
        N_STIFF = 64, N_PARAMS = 128

    type, bind(C) :: flint_init_opts
        integer(c_int) :: present = 0, method = 0
        integer(c_int) :: n_atol = 0
        type(c_ptr) :: atol = c_null_ptr
        integer(c_int) :: n_rtol = 0
        type(c_ptr) :: rtol = c_null_ptr
        integer(c_int) :: interp_on = 0, n_interp_states = 0
        type(c_ptr) :: interp_states = c_null_ptr
        real(c_double) :: min_step_size = 0, max_step_size = 0
        real(c_double) :: stepsz_params(6) = 0
        integer(c_int) :: events_on = 0, n_event_stepsz = 0
        type(c_ptr) :: event_stepsz = c_null_ptr
        integer(c_int) :: n_event_options = 0
        type(c_ptr) :: event_options = c_null_ptr
        integer(c_int) :: n_event_tol = 0
        type(c_ptr) :: event_tol = c_null_ptr
        integer(c_int) :: arith = 0
    end type
    type, bind(C) :: flint_int_opts
        integer(c_int) :: present = 0, use_const_stepsz = 0, step_advance = 0, int_steps_on = 0
        integer(c_int) :: event_mask(4) = 0
        integer(c_int) :: n_params = 0
        type(c_ptr) :: params = c_null_ptr
    end type
    type, bind(C) :: flint_steps_out
        integer(c_long) :: cap = 0
        type(c_ptr) :: xint = c_null_ptr, yint = c_null_ptr, count = c_null_ptr
    end type
    type, bind(C) :: flint_events_out
        integer(c_long) :: cap = 0
        type(c_ptr) :: rec = c_null_ptr, count = c_null_ptr
    end type
    integer, parameter, public :: FLINT_MAX_DEVICES = 16
    !> include/flint_b200.h: flint_exec, field for field (where the buffers live, which GPUs run the call, tuning knobs)
    type, bind(C), public :: flint_exec
        integer(c_int) :: mem = 0, device = 0
        type(c_ptr) :: stream = c_null_ptr
        integer(c_int) :: async = 0
        integer(c_long) :: dense_cap = 0
        integer(c_int) :: block_threads = 0, blocks_per_sm = 0, event_mode = 0, plane_variant = 0, pipeline = 0
        integer(c_int) :: round_len = 0
        integer(c_int) :: n_devices = 0                       ! > 1: the ensemble is sharded over devices(1:n_devices) inside the call
        integer(c_int) :: devices(FLINT_MAX_DEVICES) = 0
        integer(c_long) :: shard_chunk = 0
        integer(c_int) :: dense_mode = 0
        integer(c_int) :: dense_spill = 0                     ! InterpOn: k >= 2 parts whose dense output waits in host memory, 1 = automatic
    end type

    interface
        function flint_sys_create(system_id, eventset_id, n, m, sysparams, nsysparams) bind(C, name="flint_sys_create") result(h)
            import; integer(c_int), value :: system_id, eventset_id, n, m, nsysparams
            real(c_double), intent(in) :: sysparams(*); type(c_ptr) :: h
        end function
        subroutine flint_sys_destroy(h) bind(C, name="flint_sys_destroy")
            import; type(c_ptr), value :: h
        end subroutine
        function flint_erk_create() bind(C, name="flint_erk_create") result(h)
            import; type(c_ptr) :: h
        end function
        subroutine flint_erk_destroy(h) bind(C, name="flint_erk_destroy")
            import; type(c_ptr), value :: h
        end subroutine
        function flint_erk_init(h, sys, max_steps, opts) bind(C, name="flint_erk_init") result(st)
            import; type(c_ptr), value :: h, sys; integer(c_int), value :: max_steps
            type(flint_init_opts), intent(in) :: opts; integer(c_int) :: st
        end function
        function flint_erk_integrate(h, x0, y0, xf, yf, stepsz, opts, steps, events, stifftest) &
                bind(C, name="flint_erk_integrate") result(st)
            import; type(c_ptr), value :: h; real(c_double), value :: x0
            real(c_double), intent(in) :: y0(*); real(c_double), intent(inout) :: xf, stepsz
            real(c_double), intent(out) :: yf(*); type(flint_int_opts), intent(in) :: opts
            type(c_ptr), value :: steps, events, stifftest; integer(c_int) :: st
        end function
        function flint_erk_integrate_batch(h, n, x0s, x0, y0, xf, yf, stepsz, opts, counters, status, steps, events, &
                stifftest, exec) bind(C, name="flint_erk_integrate_batch") result(st)
            import; type(c_ptr), value :: h; integer(c_long), value :: n; real(c_double), value :: x0s
            type(c_ptr), value :: x0, y0, xf, yf, stepsz, counters, status, steps, events, stifftest
            type(flint_int_opts), intent(in) :: opts; type(flint_exec), intent(in) :: exec; integer(c_int) :: st
        end function
        function flint_erk_interpolate(h, xarr, g, yarr, has_save, save_coeffs) bind(C, name="flint_erk_interpolate") result(st)
            import; type(c_ptr), value :: h; integer(c_long), value :: g; integer(c_int), value :: has_save, save_coeffs
            real(c_double), intent(in) :: xarr(*); real(c_double), intent(out) :: yarr(*); integer(c_int) :: st
        end function
        function flint_erk_interpolate_batch(h, xarr, g, yarr, layout, status, has_save, save_coeffs, exec) &
                bind(C, name="flint_erk_interpolate_batch") result(st)
            import; type(c_ptr), value :: h, xarr, yarr, status; integer(c_long), value :: g
            integer(c_int), value :: layout, has_save, save_coeffs; type(flint_exec), intent(in) :: exec; integer(c_int) :: st
        end function
        function flint_erk_interpolate_batch_grids(h, xarr, g, yarr, layout, status, has_save, save_coeffs, exec) &
                bind(C, name="flint_erk_interpolate_batch_grids") result(st)
            import; type(c_ptr), value :: h, xarr, yarr, status; integer(c_long), value :: g
            integer(c_int), value :: layout, has_save, save_coeffs; type(flint_exec), intent(in) :: exec; integer(c_int) :: st
        end function
        function flint_b200_device_count() bind(C, name="flint_b200_device_count") result(n)
            import; integer(c_int) :: n
        end function
        function flint_b200_shard_count(n, n_shards, shard, chunk) bind(C, name="flint_b200_shard_count") result(c)
            import; integer(c_long), value :: n, chunk; integer(c_int), value :: n_shards, shard; integer(c_long) :: c
        end function
        function flint_erk_info(h, last_status, nsteps, naccept, nreject, nfcalls, interp_ready, x0, y0, hf, xf, yf) &
                bind(C, name="flint_erk_info") result(st)
            import; type(c_ptr), value :: h, last_status, nsteps, naccept, nreject, nfcalls, interp_ready, x0, y0, hf, xf, yf
            integer(c_int) :: st
        end function
        function flint_status_string(s) bind(C, name="flint_status_string") result(p)
            import; integer(c_int), value :: s; type(c_ptr) :: p
        end function
    end interface

    public :: flint_b200_device_count, flint_b200_shard_count

    !> DiffEqSys of the GPU build: F and G are device functors compiled into the library.
    type, public :: DiffEqSys_GPU
        integer :: n = 0, m = 0
        type(c_ptr) :: handle = c_null_ptr
    contains
        procedure :: Create => sys_create
        final :: sys_final
    end type

    !> ERK_class (src/ERK.f90:72-117)
    type, public :: ERK_class
        integer(c_int) :: status = FLINT_SUCCESS
        type(c_ptr) :: handle = c_null_ptr
        integer :: n = 0, m = 0, MaxSteps = 0
    contains
        procedure :: Init => erk_init
        procedure :: Integrate => erk_int
        procedure :: IntegrateBatch => erk_int_batch
        procedure :: Interpolate => erk_interp
        procedure :: InterpolateBatch => erk_interp_batch
        procedure :: Info => erk_info
        final :: erk_final
    end type

contains

    subroutine sys_create(me, system_id, eventset_id, sysparams)
        class(DiffEqSys_GPU), intent(inout) :: me
        integer(c_int), intent(in) :: system_id, eventset_id
        real(WP), intent(in) :: sysparams(:)
        integer, parameter :: nsys(4) = [6, 6, 3, 6], mev(0:5) = [0, 3, 2, 1, 3, 3]
        me%n = nsys(system_id); me%m = mev(eventset_id)
        me%handle = flint_sys_create(system_id, eventset_id, int(me%n, c_int), int(me%m, c_int), sysparams, size(sysparams, kind=c_int))
    end subroutine
    subroutine sys_final(me)
        type(DiffEqSys_GPU), intent(inout) :: me
        if (c_associated(me%handle)) call flint_sys_destroy(me%handle)
        me%handle = c_null_ptr
    end subroutine

    !> src/FLINT_base.f90:318-392
    subroutine erk_init(me, DE, MaxSteps, Method, ATol, RTol, InterpOn, InterpStates, MinStepSize, MaxStepSize, StepSzParams, &
                        EventsOn, EventStepSz, EventOptions, EventTol, Arith)
        class(ERK_class), intent(inout) :: me
        type(DiffEqSys_GPU), intent(in) :: DE
        integer, intent(in) :: MaxSteps
        integer(c_int), intent(in), optional :: Method, Arith
        real(WP), dimension(:), intent(in), target :: ATol, RTol          ! mandatory in practice (reference quirk Q12)
        logical, intent(in), optional :: InterpOn, EventsOn
        integer(c_int), dimension(:), intent(in), optional, target :: InterpStates, EventOptions
        real(WP), intent(in), optional :: MinStepSize, MaxStepSize
        real(WP), dimension(6), intent(in), optional :: StepSzParams
        real(WP), dimension(:), intent(in), optional, target :: EventStepSz, EventTol
        type(flint_init_opts) :: o
        if (.not. c_associated(me%handle)) me%handle = flint_erk_create()
        me%n = DE%n; me%m = DE%m; me%MaxSteps = MaxSteps
        o%n_atol = size(ATol); o%atol = c_loc(ATol); o%n_rtol = size(RTol); o%rtol = c_loc(RTol)
        if (present(Method)) then; o%present = ior(o%present, I_METHOD); o%method = Method; end if
        if (present(InterpOn)) then; o%present = ior(o%present, I_INTERP_ON); o%interp_on = merge(1, 0, InterpOn); end if
        if (present(InterpStates)) then
            o%present = ior(o%present, I_INTERP_STATES); o%n_interp_states = size(InterpStates); o%interp_states = c_loc(InterpStates)
        end if
        if (present(MinStepSize)) then; o%present = ior(o%present, I_MIN_STEP); o%min_step_size = MinStepSize; end if
        if (present(MaxStepSize)) then; o%present = ior(o%present, I_MAX_STEP); o%max_step_size = MaxStepSize; end if
        if (present(StepSzParams)) then; o%present = ior(o%present, I_STEPSZ_PARAMS); o%stepsz_params = StepSzParams; end if
        if (present(EventsOn)) then; o%present = ior(o%present, I_EVENTS_ON); o%events_on = merge(1, 0, EventsOn); end if
        if (present(EventStepSz)) then
            o%present = ior(o%present, I_EVENT_STEPSZ); o%n_event_stepsz = size(EventStepSz); o%event_stepsz = c_loc(EventStepSz)
        end if
        if (present(EventOptions)) then
            o%present = ior(o%present, I_EVENT_OPTIONS); o%n_event_options = size(EventOptions); o%event_options = c_loc(EventOptions)
        end if
        if (present(EventTol)) then
            o%present = ior(o%present, I_EVENT_TOL); o%n_event_tol = size(EventTol); o%event_tol = c_loc(EventTol)
        end if
        if (present(Arith)) then; o%present = ior(o%present, I_ARITH); o%arith = Arith; end if
        me%status = flint_erk_init(me%handle, DE%handle, int(MaxSteps, c_int), o)
    end subroutine

    subroutine fill_int_opts(o, UseConstStepSz, StepAdvance, IntStepsOn, EventMask, params)
        type(flint_int_opts), intent(out) :: o
        logical, intent(in), optional :: UseConstStepSz, StepAdvance, IntStepsOn
        logical, dimension(:), intent(in), optional :: EventMask
        real(WP), dimension(:), intent(in), optional, target :: params
        if (present(UseConstStepSz)) then; o%present = ior(o%present, N_CONST); o%use_const_stepsz = merge(1, 0, UseConstStepSz); end if
        if (present(StepAdvance)) then; o%present = ior(o%present, N_ADVANCE); o%step_advance = merge(1, 0, StepAdvance); end if
        if (present(IntStepsOn)) then; o%present = ior(o%present, N_INTSTEPS); o%int_steps_on = merge(1, 0, IntStepsOn); end if
        if (present(EventMask)) then
            o%present = ior(o%present, N_EVMASK); o%event_mask(1:size(EventMask)) = merge(1, 0, EventMask)
        end if
        if (present(params)) then; o%present = ior(o%present, N_PARAMS); o%n_params = size(params); o%params = c_loc(params); end if
    end subroutine

    !> src/FLINT_base.f90:400-468.  Xint/Yint/EventStates are allocated here (capacity MaxSteps+1 / MaxEvents) and
    !! shrunk to their true size, as the reference does (src/ERKIntegrate.f90:642-681).
    subroutine erk_int(me, X0, Y0, Xf, Yf, StepSz, UseConstStepSz, StepAdvance, IntStepsOn, Xint, Yint, EventMask, &
                       EventStates, StiffTest, params)
        class(ERK_class), intent(inout) :: me
        real(WP), intent(in) :: X0
        real(WP), dimension(me%n), intent(in) :: Y0
        real(WP), intent(inout) :: Xf
        real(WP), dimension(size(Y0)), intent(out) :: Yf
        real(WP), intent(inout) :: StepSz
        logical, intent(in), optional :: UseConstStepSz, StepAdvance, IntStepsOn
        real(WP), allocatable, dimension(:), intent(out), optional, target :: Xint
        real(WP), allocatable, dimension(:,:), intent(out), optional, target :: Yint
        logical, dimension(me%m), intent(in), optional :: EventMask
        real(WP), allocatable, dimension(:,:), intent(out), optional, target :: EventStates
        integer(c_int), intent(inout), optional, target :: StiffTest
        real(WP), dimension(:), intent(in), optional, target :: params
        type(flint_int_opts) :: o
        type(flint_steps_out), target :: so
        type(flint_events_out), target :: eo
        integer(c_int), target :: nsteps, nev
        integer, parameter :: MaxEvents = 256
        type(c_ptr) :: ps, pe, pt
        call fill_int_opts(o, UseConstStepSz, StepAdvance, IntStepsOn, EventMask, params)
        ps = c_null_ptr; pe = c_null_ptr; pt = c_null_ptr
        if (present(Xint) .and. present(Yint)) then
            allocate(Xint(me%MaxSteps + 1), Yint(me%n, me%MaxSteps + 1))
            so%cap = me%MaxSteps + 1; so%xint = c_loc(Xint); so%yint = c_loc(Yint); so%count = c_loc(nsteps); ps = c_loc(so)
        end if
        if (present(EventStates)) then
            allocate(EventStates(me%n + 2, MaxEvents))
            eo%cap = MaxEvents; eo%rec = c_loc(EventStates); eo%count = c_loc(nev); pe = c_loc(eo)
        end if
        if (present(StiffTest)) pt = c_loc(StiffTest)
        nsteps = 0; nev = 0
        me%status = flint_erk_integrate(me%handle, X0, Y0, Xf, Yf, StepSz, o, ps, pe, pt)
        if (present(Xint) .and. present(Yint)) then
            Xint = Xint(1:max(nsteps, 0)); Yint = Yint(:, 1:max(nsteps, 0))
        end if
        if (present(EventStates)) EventStates = EventStates(:, 1:min(nev, MaxEvents))
    end subroutine

    !> Ensemble form: Y0(N, n) in Fortran = SoA [n][N] in C.  Outputs likewise.  exec%n_devices > 1 shards the ensemble over
    !! the GPUs of the box inside this one call (block-cyclic chunks of 4096 trajectories, SURVEY.md 8e): the loop over
    !! trajectories of tests/test_CR3BP.f90:397-405 becomes
    !!     ex%n_devices = flint_b200_device_count();  ex%devices(1:ex%n_devices) = [(i - 1, i = 1, ex%n_devices)]
    !!     call erk%IntegrateBatch(x0, Y0, Xf, Yf, StepSz, Counters, Status, EventMask=EvMask, EventCount=nEv, EventStates=Ev, exec=ex)
    subroutine erk_int_batch(me, X0, Y0, Xf, Yf, StepSz, Counters, Status, UseConstStepSz, EventMask, EventCount, &
                             EventStates, StiffTest, exec)
        class(ERK_class), intent(inout) :: me
        real(WP), intent(in) :: X0
        real(WP), dimension(:,:), intent(in), target :: Y0                  ! (N, n)
        real(WP), dimension(:), intent(inout), target :: Xf, StepSz         ! (N)
        real(WP), dimension(:,:), intent(out), target :: Yf                 ! (N, n)
        integer(c_int), dimension(:,:), intent(out), target :: Counters     ! (N, 4)
        integer(c_int), dimension(:), intent(out), target :: Status         ! (N)
        logical, intent(in), optional :: UseConstStepSz
        logical, dimension(me%m), intent(in), optional :: EventMask
        integer(c_int), dimension(:), intent(out), optional, target :: EventCount      ! (N)
        real(WP), dimension(:,:,:), intent(out), optional, target :: EventStates       ! (N, cap, n+2)
        integer(c_int), dimension(:), intent(inout), optional, target :: StiffTest     ! (N)
        type(flint_exec), intent(in), optional :: exec
        type(flint_int_opts) :: o
        type(flint_events_out), target :: eo
        type(flint_exec) :: ex
        type(c_ptr) :: pe, pt
        call fill_int_opts(o, UseConstStepSz=UseConstStepSz, EventMask=EventMask)
        pe = c_null_ptr; pt = c_null_ptr
        if (present(EventStates) .and. present(EventCount)) then
            eo%cap = size(EventStates, 2); eo%rec = c_loc(EventStates); eo%count = c_loc(EventCount); pe = c_loc(eo)
        end if
        if (present(StiffTest)) pt = c_loc(StiffTest)
        if (present(exec)) ex = exec
        me%status = flint_erk_integrate_batch(me%handle, size(Y0, 1, kind=c_long), X0, c_null_ptr, c_loc(Y0), c_loc(Xf), c_loc(Yf), &
                                              c_loc(StepSz), o, c_loc(Counters), c_loc(Status), c_null_ptr, pe, pt, ex)
    end subroutine

    !> src/FLINT_base.f90:479-500
    subroutine erk_interp(me, Xarr, Yarr, SaveInterpCoeffs)
        class(ERK_class), intent(inout) :: me
        real(WP), dimension(1:), intent(in) :: Xarr
        real(WP), dimension(:,:), intent(out) :: Yarr
        logical, intent(in), optional :: SaveInterpCoeffs
        integer(c_int) :: hs, sv
        hs = 0; sv = 0
        if (present(SaveInterpCoeffs)) then; hs = 1; sv = merge(1, 0, SaveInterpCoeffs); end if
        me%status = flint_erk_interpolate(me%handle, Xarr, size(Xarr, kind=c_long), Yarr, hs, sv)
    end subroutine

    !> Interpolate for the ensemble of the last IntegrateBatch (src/ERKInterp.f90:28-140 once per trajectory).
    !! Xarr(G): one grid for all trajectories, Yarr(nI, G, N): trajectory i's Yarr(nI, G) is Yarr(:, :, i), exactly the array the
    !! reference's Interpolate fills (C layout 0, [N][G][nI]).  XarrPerTraj(G, N): one grid per trajectory instead.
    !! Status(N): FLINT_ERROR_INTP_ARRAY where a grid leaves its trajectory's span (:60-76).
    subroutine erk_interp_batch(me, Yarr, Xarr, XarrPerTraj, Status, SaveInterpCoeffs, exec)
        class(ERK_class), intent(inout) :: me
        real(WP), dimension(:,:,:), intent(out), target :: Yarr                       ! (nI, G, N)
        real(WP), dimension(:), intent(in), optional, target :: Xarr                   ! (G)
        real(WP), dimension(:,:), intent(in), optional, target :: XarrPerTraj          ! (G, N)
        integer(c_int), dimension(:), intent(out), optional, target :: Status          ! (N)
        logical, intent(in), optional :: SaveInterpCoeffs
        type(flint_exec), intent(in), optional :: exec
        type(flint_exec) :: ex
        type(c_ptr) :: pst
        integer(c_int) :: hs, sv
        hs = 0; sv = 0
        if (present(SaveInterpCoeffs)) then; hs = 1; sv = merge(1, 0, SaveInterpCoeffs); end if
        if (present(exec)) ex = exec
        pst = c_null_ptr
        if (present(Status)) pst = c_loc(Status)
        if (present(XarrPerTraj)) then
            me%status = flint_erk_interpolate_batch_grids(me%handle, c_loc(XarrPerTraj), size(XarrPerTraj, 1, kind=c_long), c_loc(Yarr), &
                                                          0_c_int, pst, hs, sv, ex)
        else if (present(Xarr)) then
            me%status = flint_erk_interpolate_batch(me%handle, c_loc(Xarr), size(Xarr, kind=c_long), c_loc(Yarr), 0_c_int, pst, hs, sv, ex)
        else
            me%status = FLINT_ERROR_PARAMS
        end if
    end subroutine

    !> src/FLINT_base.f90:507-533 (hf/Xf/Yf are returned in the arguments they are named after; quirk Q8 fixed)
    subroutine erk_info(me, LastStatus, StatusMsg, nSteps, nAccept, nReject, nFCalls, InterpReady, X0, Y0, hf, Xf, Yf)
        class(ERK_class), intent(inout) :: me
        integer(c_int), intent(out), target :: LastStatus
        character(len=:), allocatable, intent(out), optional :: StatusMsg
        integer(c_int), intent(out), optional, target :: nSteps, nAccept, nReject, nFCalls
        logical, intent(out), optional :: InterpReady
        real(WP), intent(out), optional, target :: X0, hf, Xf
        real(WP), dimension(:), intent(out), optional, target :: Y0, Yf
        integer(c_int), target :: ir
        integer(c_int) :: st
        character(kind=c_char), pointer :: cs(:)
        integer :: i
        type(c_ptr) :: p(10)
        p = c_null_ptr
        if (present(nSteps)) p(1) = c_loc(nSteps)
        if (present(nAccept)) p(2) = c_loc(nAccept)
        if (present(nReject)) p(3) = c_loc(nReject)
        if (present(nFCalls)) p(4) = c_loc(nFCalls)
        p(5) = c_loc(ir)
        if (present(X0)) p(6) = c_loc(X0)
        if (present(Y0)) p(7) = c_loc(Y0)
        if (present(hf)) p(8) = c_loc(hf)
        if (present(Xf)) p(9) = c_loc(Xf)
        if (present(Yf)) p(10) = c_loc(Yf)
        st = flint_erk_info(me%handle, c_loc(LastStatus), p(1), p(2), p(3), p(4), p(5), p(6), p(7), p(8), p(9), p(10))
        if (present(InterpReady)) InterpReady = ir /= 0
        if (present(StatusMsg)) then
            call c_f_pointer(flint_status_string(LastStatus), cs, [64])
            StatusMsg = ''
            do i = 1, 64
                if (cs(i) == c_null_char) exit
                StatusMsg = StatusMsg // cs(i)
            end do
        end if
    end subroutine

    subroutine erk_final(me)
        type(ERK_class), intent(inout) :: me
        if (c_associated(me%handle)) call flint_erk_destroy(me%handle)
        me%handle = c_null_ptr
    end subroutine
end module FLINT_GPU
