"""Host-side mirror of the reference interface of the hot path.

The reference's host language is C++ on Trilinos, which this image does not have; the C++ glue
that binds `FE<SC,LO,GO,NO>` to the C ABI lives in feddlib_b200/csrc/host/FE_b200.hpp (see
INTEGRATION.md).  This module mirrors the same operator interface in Python -- same method names,
argument order and meaning, and error behaviour -- so parity tests read like the reference's own
test drivers (feddlib/core/FE/tests/fe.cpp:60-101):

    domain = Domain.buildMesh(dim, FEType, N, M)          # Domain::buildMesh, Domain_def.hpp:201-263
    fe = FE(); fe.addFE(domain)                           # FE::addFE, FE_def.hpp:64-72
    A = Matrix(domain.getMapUnique(), domain.getApproxEntriesPerRow())   # Laplace_def.hpp:48
    fe.assemblyLaplace(dim, FEType, 2, A)                 # FE_def.hpp:604-667

Only the containers the hot path reads are mirrored (Map, Domain, Matrix, MultiVector-as-array).
"""
from __future__ import annotations

import numpy as np

from . import mesh as _mesh
from ._lib import BLOCK_DIAG, BLOCK_FULL, BLOCK_SCALAR, LogicError
from .engine import Context, Mesh, Pattern, assemble_div_divT_d


class Map:
    """FEDD::Map (core/LinearAlgebra/Map_decl.hpp:27-109): ordered list of global ids of one rank."""

    def __init__(self, gids, rank: int = 0, nranks: int = 1):
        self.gids = np.ascontiguousarray(gids, dtype=np.int64)
        self.rank, self.nranks = rank, nranks
        self._lookup = None

    def getNodeNumElements(self) -> int:
        return int(self.gids.size)

    def getGlobalElement(self, lid: int) -> int:
        return int(self.gids[lid])

    def getLocalElement(self, gid: int) -> int:
        if self._lookup is None:
            self._lookup = {int(g): i for i, g in enumerate(self.gids)}
        return self._lookup.get(int(gid), -1)

    def getNodeElementList(self) -> np.ndarray:
        return self.gids

    def getMaxAllGlobalIndex(self) -> int:
        return int(self.gids.max()) if self.gids.size else -1

    def buildVecFieldMap(self, numDofs: int, ordering: str = "NodeWise") -> "Map":
        """Map_def.hpp:95-108: node-wise dof numbering numDofs*g + d."""
        if ordering != "NodeWise":
            raise LogicError("Select a valid ordering: NodeWise")
        g = (self.gids[:, None] * numDofs + np.arange(numDofs, dtype=np.int64)[None, :]).ravel()
        return Map(g, self.rank, self.nranks)

    def isSameAs(self, other: "Map") -> bool:
        return np.array_equal(self.gids, other.gids)


class Domain:
    """What the hot path reads of FEDD::Domain (core/FE/Domain_decl.hpp:21-247)."""

    def __init__(self, dim, FEType, elements, pointsRepeated, mapRepeated: Map, mapUnique: Map | None = None,
                 owner=None):
        self.dim_, self.FEType_ = int(dim), FEType
        self.elementsC_ = np.ascontiguousarray(elements, dtype=np.int32)
        self.pointsRep_ = np.ascontiguousarray(pointsRepeated, dtype=np.float64)
        self.mapRepeated_ = mapRepeated
        self.owner_ = None if owner is None else np.ascontiguousarray(owner, dtype=np.int32)
        if mapUnique is None:
            if self.owner_ is None:
                mapUnique = Map(mapRepeated.gids, mapRepeated.rank, mapRepeated.nranks)
            else:  # Map::buildUniqueMap order: repeated order filtered by ownership (Map_def.hpp:201-206)
                mapUnique = Map(mapRepeated.gids[self.owner_ == mapRepeated.rank], mapRepeated.rank, mapRepeated.nranks)
        self.mapUnique_ = mapUnique

    @classmethod
    def p0_of(cls, domain: "Domain", elementMap: Map | None = None):
        """The P0 space on the elements of `domain` (pressure of assemblyDivAndDivT, FE_def.hpp:1954-1957): one pseudo-node
        per element whose global ids are the ELEMENT map (Domain::getElementMap; default: the local element numbers)."""
        ne = domain.getElementsC().shape[0]
        emap = elementMap if elementMap is not None else Map(np.arange(ne, dtype=np.int64), domain.getMapRepeated().rank,
                                                              domain.getMapRepeated().nranks)
        d = cls(domain.getDimension(), "P0", np.arange(ne, dtype=np.int32)[:, None], np.zeros((ne, domain.getDimension())), emap)
        return d

    def getElementMap(self): return self.mapRepeated_ if self.FEType_ == "P0" else None

    @classmethod
    def buildMesh(cls, dim: int, FEType: str, N: int, M: int, rank: int = 0, nranks: int = 1):
        """Built-in structured square/cube (Domain::buildMesh -> MeshStructured::buildMesh2D/3D)."""
        conn, coords, gid = _mesh.build_structured(dim, FEType, N, M, rank)
        return cls(dim, FEType, conn, coords, Map(gid, rank, nranks))

    def getDimension(self): return self.dim_
    def getFEType(self): return self.FEType_
    def getElementsC(self): return self.elementsC_
    def getPointsRepeated(self): return self.pointsRep_
    def getMapRepeated(self): return self.mapRepeated_
    def getMapUnique(self): return self.mapUnique_
    def getMapVecFieldUnique(self): return self.mapUnique_.buildVecFieldMap(self.dim_)
    def getMapVecFieldRepeated(self): return self.mapRepeated_.buildVecFieldMap(self.dim_)

    def getApproxEntriesPerRow(self) -> int:
        """Domain_def.hpp:176-197."""
        if self.dim_ == 2:
            return 20 if self.FEType_ == "P1" else 30
        return 50 if self.FEType_ == "P1" else 80


class Matrix:
    """FEDD::Matrix (core/LinearAlgebra/Matrix_def.hpp): after an assembly call it holds the local
    CSR of the fill-complete matrix on the row map it was constructed with:
    rowptr int64, colind int32 (column-map local ids), values (CUDA float64 tensor), colmap Map."""

    def __init__(self, map_: Map, numEntries: int = 0):
        self.map_ = map_
        self.numEntries = int(numEntries)
        self.rowptr = self.colind = self.values = None
        self.colmap = None
        self.domainMap = self.rangeMap = None
        self.fillComplete_ = False
        self._ctx = None

    # -- set by the engine --
    def _seat(self, ctx, rowptr, colind, values, colmap, domainMap, rangeMap, fill_complete):
        self._ctx, self.rowptr, self.colind, self.values, self.colmap = ctx, rowptr, colind, values, colmap
        self.domainMap, self.rangeMap = domainMap, rangeMap
        self.fillComplete_ = fill_complete

    def getMap(self): return self.map_
    def isFillComplete(self): return self.fillComplete_
    def getNodeNumRows(self): return self.map_.getNodeNumElements()

    def resumeFill(self):
        self.fillComplete_ = False

    def fillComplete(self, domainMap: Map | None = None, rangeMap: Map | None = None):
        """Matrix_def.hpp:192-199.  The engine already delivers merged, sorted CSR rows; this only
        records the domain/range maps (callers do resumeFill -> scale -> fillComplete(dom, rng))."""
        if domainMap is not None:
            self.domainMap, self.rangeMap = domainMap, rangeMap
        self.fillComplete_ = True

    def scale(self, alpha: float):
        """Matrix_def.hpp:257 (Matrix::scale)."""
        self._ctx.scale_d(self.values, alpha)

    def numpy_values(self) -> np.ndarray:
        self._ctx.synchronize()
        return self.values.cpu().numpy()

    def toScipy(self):
        import scipy.sparse as sp
        n_rows = self.map_.getNodeNumElements()
        ncols = self.colmap.getNodeNumElements()
        nnz = int(self.rowptr[n_rows])
        return sp.csr_matrix((self.numpy_values()[:nnz], self.colind[:nnz], self.rowptr[: n_rows + 1]),
                             shape=(n_rows, ncols))


class FE:
    """FE<SC,LO,GO,NO> (core/FE/FE_decl.hpp:40-488) -- the seven hot-path entry points."""

    def __init__(self, saveAssembly: bool = False, device: int = 0, ctx: Context | None = None):
        self.domainVec_ = []
        self.saveAssembly_ = saveAssembly
        self.ctx = ctx or Context(device)
        self._meshes = {}    # id(domain) -> Mesh
        self._patterns = {}  # (id(row domain), id(col domain)) -> Pattern
        self._expanded = {}  # (pattern key, rd, cd, mode) -> (rowptr, colind)

    # FE_def.hpp:64-72 -- the natural place for the one-time upload
    def addFE(self, domain: Domain):
        self.domainVec_.append(domain)
        self._meshes[id(domain)] = Mesh(self.ctx, domain.getDimension(), domain.getElementsC(),
                                        domain.getPointsRepeated())

    def setScatterMode(self, mode):
        self.ctx.set_scatter_mode(mode)

    # FE_def.hpp:6932-6953
    def checkFE(self, dim: int, FEType: str) -> int:
        found = -1
        for i, d in enumerate(self.domainVec_):
            if d.getDimension() == dim and d.getFEType() == FEType:
                found = i
        if found < 0:
            raise LogicError("Combination of dimenson(2/3) and FE Type(P1/P2) not defined yet. Use addFE(domain)")
        return found

    # ---- internals ----
    def _pattern(self, drow: Domain, dcol: Domain) -> Pattern:
        key = (id(drow), id(dcol))
        if key not in self._patterns:
            self._patterns[key] = Pattern(self.ctx, self._meshes[id(drow)], self._meshes[id(dcol)])
        return self._patterns[key]

    def _csr(self, pat: Pattern, key, rd, cd, mode):
        k = (key, rd, cd, mode)
        if k not in self._expanded:
            self._expanded[k] = pat.expand(rd, cd, mode)
        return self._expanded[k]

    def _finish(self, A: Matrix, drow: Domain, dcol: Domain, pat: Pattern, values, rd, cd, mode, callFillComplete,
                domainMap=None, rangeMap=None):
        rowmap = drow.getMapUnique() if rd == 1 else drow.getMapUnique().buildVecFieldMap(rd)
        if A.getMap().getNodeNumElements() != rowmap.getNodeNumElements():
            raise LogicError("Matrix row map does not match the unique map of the FE space "
                             f"({A.getMap().getNodeNumElements()} rows given, {rowmap.getNodeNumElements()} needed)")
        colmap = dcol.getMapUnique() if cd == 1 else dcol.getMapUnique().buildVecFieldMap(cd)
        rowptr, colind = self._csr(pat, (id(drow), id(dcol)), rd, cd, mode)
        A._seat(self.ctx, rowptr, colind, values, colmap, domainMap or colmap, rangeMap or rowmap, callFillComplete)

    # ---- FE_def.hpp:4694-4766 ----
    def assemblyRHS(self, dim, FEType, a, fieldType, func, funcParameter):
        """`a`: repeated (vector-field) array, added to in place; `func(x, res, parameters)` fills res[0:dim] and is
        evaluated once, as in the reference (:4735); funcParameter[-1] is the degree of the function (:4716)."""
        if FEType == "P0":
            raise LogicError("Not implemented for P0")
        if a is None:
            raise RuntimeError("MultiVector in assemblyConstRHS is null.")
        if fieldType not in ("Scalar", "Vector"):
            raise LogicError("Invalid field type.")
        d = self.domainVec_[self.checkFE(dim, FEType)]
        res = np.zeros(dim)
        func(np.zeros(dim), res, np.asarray(funcParameter, dtype=np.float64))
        degFunc = int(funcParameter[-1] + 1.e-14)
        pat = self._pattern(d, d)
        a += pat.assemble_rhs(res, degFunc, fieldType == "Vector")

    # ---- FE_def.hpp:2407-2735 ----
    def assemblyStress(self, dim, FEType, A: Matrix, func, parameters=None, callFillComplete=True):
        """func(x, parameters) -> float is the reference's CoeffFunc_Type; it is evaluated on the host at the physical
        quadrature points of every element, exactly where the reference evaluates it."""
        if FEType == "P0":
            raise LogicError("Not implemented for P0")
        d = self.domainVec_[self.checkFE(dim, FEType)]
        pat = self._pattern(d, d)
        xyz = pat.stress_points(d.elementsC_, d.pointsRep_)
        f = np.array([[func(x, parameters) for x in el] for el in xyz], dtype=np.float64)
        values = self.ctx.empty_values(pat.nnz(dim, dim, BLOCK_FULL))
        if f.size == 0 or np.all(f == f.flat[0]):
            pat.assemble_stress_d(values, float(f.flat[0]) if f.size else 1.0)
        else:
            import torch
            pat.assemble_stress_d(values, torch.from_numpy(f.ravel()).to(values.device))
        self._finish(A, d, d, pat, values, dim, dim, BLOCK_FULL, callFillComplete)

    # ---- FE_def.hpp:2151-2220 ----
    def assemblyBDStabilization(self, dim, FEType, A: Matrix, callFillComplete=True):
        if FEType != "P1":
            raise LogicError("Only implemented for P1. Q1 is equivalent but we need to adjust scaling for the reference element.")
        d = self.domainVec_[self.checkFE(dim, FEType)]
        pat = self._pattern(d, d)
        values = self.ctx.empty_values(pat.nnz())
        pat.assemble_bdstab_d(values)
        self._finish(A, d, d, pat, values, 1, 1, BLOCK_SCALAR, callFillComplete)

    # ---- FE_def.hpp:454-521 ----
    def assemblyMass(self, dim, FEType, fieldType, A: Matrix, callFillComplete=True):
        if FEType == "P0":
            raise LogicError("Not implemented for P0")
        d = self.domainVec_[self.checkFE(dim, FEType)]
        if fieldType not in ("Scalar", "Vector"):
            raise LogicError("Specify valid vieldType for assembly of mass matrix.")
        vec = fieldType == "Vector"
        pat = self._pattern(d, d)
        values = self.ctx.empty_values(pat.nnz(dim, dim, BLOCK_DIAG) if vec else pat.nnz())
        pat.assemble_mass_d(values, vec_field=vec)
        if vec:
            self._finish(A, d, d, pat, values, dim, dim, BLOCK_DIAG, callFillComplete)
        else:
            self._finish(A, d, d, pat, values, 1, 1, BLOCK_SCALAR, callFillComplete)

    # ---- FE_def.hpp:604-667 ----
    def assemblyLaplace(self, dim, FEType, degree, A: Matrix, callFillComplete=True, FELocExternal=-1):
        if FEType == "P0":
            raise LogicError("Not implemented for P0")
        FEloc = self.checkFE(dim, FEType) if FELocExternal < 0 else FELocExternal
        d = self.domainVec_[FEloc]
        pat = self._pattern(d, d)
        values = self.ctx.empty_values(pat.nnz())
        pat.assemble_laplace_d(values, vec_field=False)
        self._finish(A, d, d, pat, values, 1, 1, BLOCK_SCALAR, callFillComplete)

    # ---- FE_def.hpp:670-734 ----
    def assemblyLaplaceVecField(self, dim, FEType, degree, A: Matrix, callFillComplete=True):
        if FEType in ("P1-disc", "P0"):
            raise LogicError("Not implemented for P0 or P1-disc")
        d = self.domainVec_[self.checkFE(dim, FEType)]
        pat = self._pattern(d, d)
        values = self.ctx.empty_values(pat.nnz(dim, dim, BLOCK_DIAG))
        pat.assemble_laplace_d(values, vec_field=True)
        self._finish(A, d, d, pat, values, dim, dim, BLOCK_DIAG, callFillComplete)

    # ---- FE_def.hpp:2739-3040 ----
    def assemblyLinElasXDim(self, dim, FEType, A: Matrix, lambda_, mu, callFillComplete=True):
        if FEType == "P0":
            raise LogicError("Not implemented for P0")
        d = self.domainVec_[self.checkFE(dim, FEType)]
        pat = self._pattern(d, d)
        values = self.ctx.empty_values(pat.nnz(dim, dim, BLOCK_FULL))
        pat.assemble_linelas_d(values, lambda_, mu)
        self._finish(A, d, d, pat, values, dim, dim, BLOCK_FULL, callFillComplete)

    def _velocity(self, d: Domain, u):
        import torch
        u = np.ascontiguousarray(u, dtype=np.float64).ravel()
        if u.size != d.getDimension() * d.getMapRepeated().getNodeNumElements():
            raise LogicError("velocity vector must live on the repeated vector-field map (dim * #repeated nodes)")
        return torch.from_numpy(u).to(f"cuda:{self.ctx.device}")

    # ---- FE_def.hpp:1685-1836 ----
    def assemblyAdvectionVecField(self, dim, FEType, A: Matrix, u, callFillComplete=True):
        if getattr(u, "ndim", 1) > 1 and u.shape[1] > 1:
            raise LogicError("Implement for numberMV > 1 .")
        if FEType == "P0":
            raise LogicError("Not implemented for P0")
        d = self.domainVec_[self.checkFE(dim, FEType)]
        pat = self._pattern(d, d)
        values = self.ctx.empty_values(pat.nnz(dim, dim, BLOCK_DIAG))
        pat.assemble_advection_d(values, self._velocity(d, u))
        self._finish(A, d, d, pat, values, dim, dim, BLOCK_DIAG, callFillComplete)

    # ---- FE_def.hpp:1839-1929 ----
    def assemblyAdvectionInUVecField(self, dim, FEType, A: Matrix, u, callFillComplete=True):
        if getattr(u, "ndim", 1) > 1 and u.shape[1] > 1:
            raise LogicError("Implement for numberMV > 1 .")
        if FEType == "P0":
            raise LogicError("Not implemented for P0")
        d = self.domainVec_[self.checkFE(dim, FEType)]
        pat = self._pattern(d, d)
        values = self.ctx.empty_values(pat.nnz(dim, dim, BLOCK_FULL))
        pat.assemble_advection_in_u_d(values, self._velocity(d, u))
        self._finish(A, d, d, pat, values, dim, dim, BLOCK_FULL, callFillComplete)

    # ---- FE_def.hpp:1932-2057 (and the "Fast" variant :2061-2148, same result) ----
    def assemblyDivAndDivT(self, dim, FEType1, FEType2, degree, Bmat: Matrix, BTmat: Matrix, map1: Map, map2: Map,
                           callFillComplete=True):
        if FEType2 in ("P1-disc", "P1-disc-global"):
            raise LogicError("Not implemented for P1-disc pressure on the B200 path (the reference pairs it with Q2 meshes)")
        if FEType2 == "P0" and dim != 2:
            # FE::phi has a P0 case for dim 1 and 2 only (FE_def.hpp:4955, 4993): in 3D the reference reads an uninitialised value
            raise LogicError("P0 pressure is implemented for dim == 2 only (as in the reference)")
        dv = self.domainVec_[self.checkFE(dim, FEType1)]
        dp = self.domainVec_[self.checkFE(dim, FEType2)]
        if dv.getElementsC().shape[0] != dp.getElementsC().shape[0]:
            raise LogicError("velocity and pressure domains must share the element list")
        patB, patBT = self._pattern(dp, dv), self._pattern(dv, dp)
        vB = self.ctx.empty_values(patB.nnz(1, dim, BLOCK_FULL))
        vBT = self.ctx.empty_values(patBT.nnz(dim, 1, BLOCK_FULL))
        assemble_div_divT_d(self.ctx, patB, patBT, vB, vBT)
        # Bmat->fillComplete(map1, map2); BTmat->fillComplete(map2, map1)   (FE_def.hpp:2052-2055)
        self._finish(Bmat, dp, dv, patB, vB, 1, dim, BLOCK_FULL, callFillComplete, map1, map2)
        self._finish(BTmat, dv, dp, patBT, vBT, dim, 1, BLOCK_FULL, callFillComplete, map2, map1)

    assemblyDivAndDivTFast = assemblyDivAndDivT

    # ---- north_star aliases ----
    def assemblyElasticity(self, dim, FEType, A, lambda_, mu, callFillComplete=True):
        self.assemblyLinElasXDim(dim, FEType, A, lambda_, mu, callFillComplete)

    def assemblyStokes(self, dim, FETypeVelocity, FETypePressure, A, Bmat, BTmat, callFillComplete=True):
        """Stokes::assemble (problems/specific/Stokes_def.hpp:47-95): vector Laplace + B + B^T."""
        self.assemblyLaplaceVecField(dim, FETypeVelocity, 2, A, callFillComplete)
        dv = self.domainVec_[self.checkFE(dim, FETypeVelocity)]
        dp = self.domainVec_[self.checkFE(dim, FETypePressure)]
        self.assemblyDivAndDivT(dim, FETypeVelocity, FETypePressure, 2, Bmat, BTmat, dv.getMapVecFieldUnique(),
                                dp.getMapUnique(), callFillComplete)

    def assemblyNavierStokesJacobian(self, dim, FEType, A: Matrix, u, rho, nu, newton=True, callFillComplete=True):
        """Fused (0,0) block rho*nu*A + rho*N(u) [+ rho*W(u)] of NavierStokes::reAssemble
        (problems/specific/NavierStokes_def.hpp:140-152, 282-322) on the union (full-block) pattern."""
        d = self.domainVec_[self.checkFE(dim, FEType)]
        pat = self._pattern(d, d)
        values = self.ctx.empty_values(pat.nnz(dim, dim, BLOCK_FULL))
        pat.assemble_ns_jacobian_d(values, self._velocity(d, u), rho, nu, newton)
        self._finish(A, d, d, pat, values, dim, dim, BLOCK_FULL, callFillComplete)
