"""Unit tests of the Fortran-subset executor (tests/golden/f90exec.py) on small programs written for this purpose:
the semantics the reference vectors depend on (kinds and promotion, integer division, powers, assignment conversion,
array sections / sequence association / unchecked bounds, module state, use association, control flow), plus a live
run of the reference's own sources against the oracle when /root/reference is present (it is absent on the GPU box).
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import f90exec as fx  # noqa: E402

SRC = """
module kinds
    integer, parameter:: double=selected_real_kind(13,200)
end module kinds

module state
    use kinds
    implicit none
    real(kind=double), parameter, public:: pi=3.1415926535897932384626433d0
    integer, public, save:: counter
    real(kind=double), dimension(:,:), allocatable, public:: tab
    integer, dimension(:), allocatable, public:: iv
    contains
    subroutine bump(n)
        integer, intent(in):: n
        counter=counter+n
    end subroutine bump
end module state

module work
    use kinds
    use state, only: counter, bump, tab, pi, iv
    implicit none
    contains
    integer function idiv(a,b)
        integer, intent(in):: a,b
        idiv=a/b
    end function idiv

    real(kind=double) function single_literal()
        ! 0.1 is a default-real literal: rounded to float32 before it is widened (SURVEY Q1)
        single_literal=0.1
    end function single_literal

    real(kind=double) function mixed(x)
        real(kind=double), intent(in):: x
        real:: s
        s=1.1
        mixed=x*s+2*x/3          ! double*single -> double ; 2*x/3 is (2*x)/3 in double
    end function mixed

    complex(kind=double) function cm(x,y)
        real(kind=double), intent(in):: x,y
        cm=cmplx(x,y)            ! no kind: single-precision complex (SURVEY Q2)
    end function cm

    complex(kind=double) function cmd(x,y)
        real(kind=double), intent(in):: x,y
        cmd=cmplx(x,y,kind=double)
    end function cmd

    real(kind=double) function repart(x,y)
        real(kind=double), intent(in):: x,y
        complex(kind=double):: z
        z=cmplx(x,y,kind=double)
        repart=z*z               ! complex -> real assignment keeps the real part
    end function repart

    integer function trunc(x)
        real(kind=double), intent(in):: x
        trunc=x                  ! real -> integer assignment truncates toward zero
    end function trunc

    real(kind=double) function powers(x)
        real(kind=double), intent(in):: x
        powers=x**3-2**3+(-x)**2
    end function powers

    subroutine swap(a,b)
        integer:: a,b,t
        t=a; a=b; b=t
    end subroutine swap

    integer function use_swap(i,j)
        integer, intent(in):: i,j
        integer:: p,q
        p=i; q=j
        call swap(p,q)
        use_swap=p*100+q
    end function use_swap

    subroutine fill(v,n)
        integer, intent(in):: n
        integer, dimension(n), intent(out):: v
        integer:: k
        do k=1,n
            v(k)=k*k
        end do
    end subroutine fill

    integer function sections()
        integer, dimension(3,4):: m
        integer, dimension(6):: w
        integer:: i,j
        do i=1,3
            do j=1,4
                m(i,j)=10*i+j
            end do
        end do
        w=0
        call fill(w(3),3)        ! sequence association: the dummy starts at w(3)
        sections=sum(m(2,:))+sum(m(:,3))+w(3)+w(4)+w(5)+w(6)+maxval(m(1:2,2:3))
    end function sections

    integer function beyond(v,n)
        integer, intent(in):: n
        integer, dimension(n), intent(in):: v
        beyond=v(n+2)            ! Fortran does not check bounds: reads the caller's storage (merge_sort relies on it)
    end function beyond

    integer function call_beyond()
        integer, dimension(8):: w
        w=(/(10*k,k=1,8)/)
        call_beyond=beyond(w,3)
    end function call_beyond

    integer function loops()
        integer:: i,acc
        acc=0
        do i=1,10
            if (mod(i,2).eq.0) cycle
            if (i.gt.7) exit
            acc=acc+i
        end do
        loops=acc*100+i          ! 1+3+5+7 = 16, exit at i=9
        do i=5,1,-2
            acc=acc+1
        end do
        loops=loops*10+i         ! after a completed loop the variable is one step past the end: -1
    end function loops

    integer function cases(k)
        integer, intent(in):: k
        select case(k)
            case(1,3)
                cases=13
            case(4:6)
                cases=46
            case default
                cases=-1
        end select
    end function cases

    integer function modstate()
        call bump(2); call bump(5)
        if (.not.allocated(tab)) allocate(tab(2,3))
        tab=1.5d0
        tab(2,:)=(/1.d0,2.d0,3.d0/)
        iv=(/4,5,6,7/)           ! assignment to an unallocated allocatable allocates it (Fortran 2003)
        modstate=counter*1000+int(sum(tab))*10+size(iv)
    end function modstate

    complex(kind=double) function cdivide(a,b)
        complex(kind=double), intent(in):: a,b
        cdivide=a/b
    end function cdivide
end module work
"""


@pytest.fixture(scope="module")
def rt(tmp_path_factory):
    p = tmp_path_factory.mktemp("f90") / "unit.f90"
    p.write_text(SRC)
    return fx.Runtime([str(p)])


def test_integer_division_truncates_toward_zero(rt):
    assert [rt.call("work", "idiv", a, b) for a, b in ((7, 2), (-7, 2), (7, -2), (1, 3))] == [3, -3, -3, 0]


def test_default_real_literals_and_mixed_kind_promotion(rt):
    assert rt.call("work", "single_literal") == np.float64(np.float32(0.1)) != 0.1
    x = np.float64(0.3)
    assert rt.call("work", "mixed", x) == x * np.float64(np.float32(1.1)) + (2 * x) / 3


def test_cmplx_without_kind_is_single_precision(rt):
    z = rt.call("work", "cm", 0.1, 0.7)
    assert z == complex(np.float32(0.1), np.float32(0.7)) and type(z) is np.complex128
    assert rt.call("work", "cmd", 0.1, 0.7) == complex(0.1, 0.7)


def test_assignment_converts_to_the_declared_type(rt):
    assert rt.call("work", "repart", 3.0, 2.0) == 5.0                 # Re[(3+2i)^2]
    assert [rt.call("work", "trunc", v) for v in (2.9, -2.9)] == [2, -2]


def test_integer_powers_are_multiplication_chains(rt):
    x = np.float64(1.1)
    assert rt.call("work", "powers", x) == (x * x) * x - 8 + (-x) * (-x)


def test_scalar_arguments_are_passed_by_reference(rt):
    assert rt.call("work", "use_swap", 3, 4) == 403


def test_sections_sequence_association_and_unchecked_bounds(rt):
    # sum(m(2,:)) = 21+22+23+24 = 90, sum(m(:,3)) = 13+23+33 = 69, w(3:5) = 1,4,9, w(6) = 0, maxval(m(1:2,2:3)) = 23
    assert rt.call("work", "sections") == 90 + 69 + 14 + 23
    assert rt.call("work", "call_beyond") == 50


def test_control_flow(rt):
    assert rt.call("work", "loops") == (16 * 100 + 9) * 10 - 1
    assert [rt.call("work", "cases", k) for k in (1, 3, 5, 7)] == [13, 13, 46, -1]


def test_module_state_and_allocatables(rt):
    rt.mod("state").counter = 0
    assert rt.call("work", "modstate") == 7 * 1000 + int(1.5 * 3 + 6.0) * 10 + 4
    assert rt.mod("state").tab.a.flags["F_CONTIGUOUS"] and rt.mod("state").pi == np.float64(3.141592653589793)


def test_complex_division_follows_gfortran(rt):
    a, b = complex(1.0, 2.0), complex(3.0, -0.5)
    z = rt.call("work", "cdivide", a, b)
    ratio = b.imag / b.real
    div = b.imag * ratio + b.real
    assert z == complex((a.imag * ratio + a.real) / div, (a.imag - a.real * ratio) / div)     # Smith, |br| >= |bi|


@pytest.mark.skipif(not os.path.isdir("/root/reference/MoVFEM_3DMT/src"), reason="reference sources not present")
@pytest.mark.parametrize("mn,dirichlet,sch", [(8, 0, 0), (8, 1, 1)])
def test_live_execution_of_the_reference_matches_the_oracle(mn, dirichlet, sch):
    """Runs the reference's own Fortran (global_vfem etc.) right now on a 2x2x3 mesh and compares with the oracle: the
    committed tests/golden/ref_*.npz are reproducible, not hand-made."""
    import ref_exec
    from movfem_b200 import mesh
    from oracle.oracle import Oracle
    m = mesh.build_model("live", 2, 2, mn, 1000., 1100., 900., 1, 1, 0, dirichlet=dirichlet, gpml_sch=sch, freqs=(0.5,),
                         sigma_fn=mesh._layered((600., 600., 0., 900.)), topo_amp=30.0)
    r = ref_exec.ReferenceRun(m)
    o = Oracle(m)
    assert (r.nne, r.nnze) == (o.nne, o.nnze) and np.array_equal(r.gne, o.gne())
    ref = r.frequency(1)
    res = o.assemble(m.omega(1), m.sigma_for(1), faithful=True)
    assert np.array_equal(ref["irn"], res["irn"]) and np.array_equal(ref["jcn"], res["jcn"])
    assert np.array_equal(ref["a"], res["a"]) and np.array_equal(ref["rhs"], res["rhs"])      # bit for bit


@pytest.mark.skipif(not os.path.isdir("/root/reference/MoVFEM_3DMT/src"), reason="reference sources not present")
def test_live_update_sigma_q12_matches_the_mesh_model():
    """geometry.f90:144-153 update_sigma indexes g_sigma(6,npt) as g_sigma(i<=npt, j<=6): executed with
    storage-sequence addressing, as compiled Fortran behaves, it must leave exactly what mesh.Model.sigma_for models --
    the g_sigma every multi-frequency test and the frequency-sharded runs hand to the assembly."""
    from movfem_b200 import mesh
    src = "/root/reference/MoVFEM_3DMT/src/"
    rt = fx.Runtime([src + "kind_param.f90", src + "geometry.f90"], linear=("g_sigma",))
    m = mesh.build_model("q12", 3, 3, 20, 1000., 1100., 900., 1, 1, 1, freqs=(0.5, 2.0, 7.0), sigma_fn=mesh._layered((600., 600., 0., 900.)))
    g = rt.mod("geometry")
    g.g_npt = m.npt
    g.g_freq.a = np.array(m.freqs)
    g.g_sigma.a = np.asfortranarray(m.sigma_initial().T.copy())
    for ii in (1, 2, 3):
        rt.call("geometry", "update_omega", ii)
        rt.call("geometry", "update_sigma")
        assert g.omega == m.omega(ii)
        assert np.array_equal(g.g_sigma.a.T, m.sigma_for(ii))
    assert not np.array_equal(m.sigma_for(1), m.sigma_for(3))
    # only the first npt+30 storage positions are ever refreshed: the tail keeps the first frequency's imaginary part
    tail = g.g_sigma.a.T.reshape(-1)[m.npt + 30:]
    assert np.all(tail.imag[tail.imag != 0] == np.float32(mesh.EPS0 * m.omega(1)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/MoVFEM_3DMT/src"), reason="reference sources not present")
def test_live_extension_reproduces_the_config1_grid_lines():
    """geometry.f90:223-343 `extension` executed with the shipped example's parameters (100 km x 100 km inner zone,
    dx = dy = 1990 m, dz = 2000 m, sigma_bg = 0.01 S/m, f = 0.1 Hz): nextd = 4, 59 x 59 grid lines, extension cells
    1.3*i*dx -- the shape mesh.config(1) (BASELINE configs[0]) is built from.  The reference centres the lines on the
    extended *input* range, the stand-in on the lines themselves: equal up to a translation."""
    from movfem_b200 import mesh
    src = "/root/reference/MoVFEM_3DMT/src/"
    rt = fx.Runtime([src + "kind_param.f90", src + "geometry.f90"])
    g = rt.mod("geometry")
    g.g_nf = 1
    g.g_freq.a = np.array([0.1])
    g.xmin, g.xmax, g.ymin, g.ymax = np.float64(0.0), np.float64(1.0e5), np.float64(0.0), np.float64(1.0e5)
    g.zmin, g.zmax = np.float64(-50000.0), np.float64(20000.0)
    nsf, nsp = 4, np.array([4, 4, 4, 4], dtype=np.int64)
    xto, yto, zto = (np.zeros((nsf, 8), order="F") for _ in range(3))
    rt.call("geometry", "extension", 0.01, 1990.0, 1990.0, 2000.0, nsf, nsp, xto, yto, zto)
    m = mesh.config(1)
    assert (g.nextd, g.g_nx, g.g_ny) == (m.nextd, m.g_nx, m.g_ny) == (4, 59, 59)
    np.testing.assert_allclose(np.diff(g.x.a), np.diff(m.g_xp), rtol=1e-12)
    np.testing.assert_allclose(np.diff(g.y.a), np.diff(m.g_yp), rtol=1e-12)
    np.testing.assert_allclose(g.dmz.a, 1.3 * 2000.0 * np.arange(1, 5), rtol=1e-15)       # z extension: 26 km per side
