!> ISO_C_BINDING interface to libpfmds_b200.so (include/pfmds_b200.h).
!> A maintainer of PFMDS adds this module to the build (after md_general in unix_bash_compile.sh) and
!> replaces the body of the `do md_step` loop of md() (MOLECULAR_DYNAMICS/md_simulation.f90:114-243) by
!> the calls shown in INTEGRATION.md.  Built with -fdefault-real-8, `real` below is real(c_double).
!> NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Fortran compiler is available): source only.
module pfmds_b200
use iso_c_binding
implicit none

integer(c_int), parameter :: PFMDS_NVE = 0, PFMDS_NVT = 1, PFMDS_NVMS = 2

interface
	integer(c_int) function pfmds_create(ctx,device,n_atoms,positions,velocities,masses,box_size) bind(C,name='pfmds_create')
		import; type(c_ptr) :: ctx; integer(c_int),value :: device,n_atoms
		real(c_double) :: positions(3,*),velocities(3,*),masses(*),box_size(3)
	end function
	integer(c_int) function pfmds_set_group(ctx,group_num,n,indexes) bind(C,name='pfmds_set_group')
		import; type(c_ptr),value :: ctx; integer(c_int),value :: group_num,n; integer(c_int) :: indexes(*)
	end function
	integer(c_int) function pfmds_set_roles(ctx,all_moving,xyz_moving,z_moving,all_atoms) bind(C,name='pfmds_set_roles')
		import; type(c_ptr),value :: ctx; integer(c_int),value :: all_moving,xyz_moving,z_moving,all_atoms
	end function
	integer(c_int) function pfmds_add_group_change(ctx,group_from,group_to,change_ts1,change_ts2,change_frec) bind(C,name='pfmds_add_group_change')
		import; type(c_ptr),value :: ctx; integer(c_int),value :: group_from,group_to,change_ts1,change_ts2,change_frec
	end function
	integer(c_int) function pfmds_group_size(ctx,group_num,n) bind(C,name='pfmds_group_size')
		import; type(c_ptr),value :: ctx; integer(c_int),value :: group_num; integer(c_int) :: n
	end function
	integer(c_int) function pfmds_add_nhc(ctx,group_num,temperature,M,q1) bind(C,name='pfmds_add_nhc')
		import; type(c_ptr),value :: ctx; integer(c_int),value :: group_num,M; real(c_double),value :: temperature,q1
	end function
	integer(c_int) function pfmds_set_misc(ctx,zero_momentum_period,invert_z_vel) bind(C,name='pfmds_set_misc')
		import; type(c_ptr),value :: ctx; integer(c_int),value :: zero_momentum_period,invert_z_vel
	end function
	integer(c_int) function pfmds_add_interaction(ctx,name,n_params,params,nl_n,group_nums,neighb_num_max,r_cut,update_period) &
	bind(C,name='pfmds_add_interaction')
		import; type(c_ptr),value :: ctx; character(kind=c_char) :: name(*); integer(c_int),value :: n_params,nl_n
		real(c_double) :: params(*),r_cut(*); integer(c_int) :: group_nums(*),neighb_num_max(*),update_period(*)
	end function
	integer(c_int) function pfmds_advance(ctx,integrator,dt,first_md_step,n_steps) bind(C,name='pfmds_advance')
		import; type(c_ptr),value :: ctx; integer(c_int),value :: integrator,first_md_step,n_steps; real(c_double),value :: dt
	end function
	integer(c_int) function pfmds_advance_with_energy(ctx,integrator,dt,first_md_step,n_steps) bind(C,name='pfmds_advance_with_energy')
		import; type(c_ptr),value :: ctx; integer(c_int),value :: integrator,first_md_step,n_steps; real(c_double),value :: dt
	end function
	integer(c_int) function pfmds_advance_logged(ctx,integrator,dt,first_md_step,n_steps,log_period,rows,row_len,n_rows) &
	bind(C,name='pfmds_advance_logged')
		import; type(c_ptr),value :: ctx; integer(c_int),value :: integrator,first_md_step,n_steps,log_period,row_len
		real(c_double),value :: dt; real(c_double) :: rows(row_len,*); integer(c_int) :: n_rows
	end function
	integer(c_int) function pfmds_state_size(ctx,n_doubles) bind(C,name='pfmds_state_size')
		import; type(c_ptr),value :: ctx; integer(c_long_long) :: n_doubles
	end function
	integer(c_int) function pfmds_save_state(ctx,blob) bind(C,name='pfmds_save_state')
		import; type(c_ptr),value :: ctx; real(c_double) :: blob(*)
	end function
	integer(c_int) function pfmds_restore_state(ctx,positions,velocities,blob) bind(C,name='pfmds_restore_state')
		import; type(c_ptr),value :: ctx; real(c_double) :: positions(3,*),velocities(3,*),blob(*)
	end function
	integer(c_int) function pfmds_energies(ctx,e_inter,kinetic_energy,temperature,e_nhc) bind(C,name='pfmds_energies')
		import; type(c_ptr),value :: ctx; real(c_double) :: e_inter(*),kinetic_energy,temperature,e_nhc(*)
	end function
	integer(c_int) function pfmds_diagnostics(ctx,fs,mc,mcv,max_velocity,nl_load) bind(C,name='pfmds_diagnostics')
		import; type(c_ptr),value :: ctx; real(c_double) :: fs(3),mc(3),mcv(3),max_velocity; integer(c_int) :: nl_load(*)
	end function
	integer(c_int) function pfmds_download(ctx,positions,velocities,forces) bind(C,name='pfmds_download')
		import; type(c_ptr),value :: ctx; real(c_double) :: positions(3,*),velocities(3,*),forces(3,*)
	end function
	integer(c_int) function pfmds_synchronize(ctx) bind(C,name='pfmds_synchronize')
		import; type(c_ptr),value :: ctx
	end function
	type(c_ptr) function pfmds_last_error(ctx) bind(C,name='pfmds_last_error')
		import; type(c_ptr),value :: ctx
	end function
	integer(c_int) function pfmds_destroy(ctx) bind(C,name='pfmds_destroy')
		import; type(c_ptr),value :: ctx
	end function
end interface

contains

!> Stops with the library's message (the reference's own `stop` text where one exists).
subroutine pfmds_check(ctx,rc)
	type(c_ptr) :: ctx
	integer(c_int) :: rc
	character(kind=c_char),pointer :: msg(:)
	integer :: k
	if (rc/=0) then
		call c_f_pointer(pfmds_last_error(ctx),msg,[512])
		k = 1
		do while (k<512 .and. msg(k)/=c_null_char); k = k+1; enddo
		write(*,*) msg(1:k-1)
		stop
	endif
end subroutine pfmds_check

end module pfmds_b200
