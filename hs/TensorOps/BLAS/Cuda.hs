{-# LANGUAGE DataKinds                #-}
{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE GADTs                    #-}
{-# LANGUAGE KindSignatures           #-}
{-# LANGUAGE LambdaCase               #-}
{-# LANGUAGE ScopedTypeVariables      #-}
{-# LANGUAGE TypeFamilies             #-}

-- | TensorOps.BLAS.Cuda — the reference-side binding of libtops_b200.so.
--
-- Drop this file into the reference tree as @src/TensorOps/BLAS/Cuda.hs@, add it to @exposed-modules@ and
-- @extra-libraries: tops_b200@ in @tensor-ops.cabal@, and select the backend by type application exactly as the
-- apps select hmatrix today (@Proxy \@(BTensorV CuMat)@ instead of @Proxy \@(BTensorV (HMat Double))@,
-- app/Dots.hs:145).  `Tensor (BTensor v CuMat)` then comes for free from Backend/BTensor.hs:775-786.
--
-- The flat-storage `instance Tensor CuTensor` (rank >= 3 without nested boxed vectors, rank-0 results kept on the device) is the
-- sibling module TensorOps.Backend.Cuda; this one is the `BLAS` dictionary for users who want to keep `BTensor`.
--
-- STATUS: written against include/tops_b200.h; NOT COMPILED — there is no GHC in the build image (SURVEY.md fact 5).
-- It is deliberately thin and mechanical: every method is one foreign call; all arithmetic, shape checks and error
-- reporting live behind the C ABI, which is what tests/ exercises.
--
-- Two things differ from `instance BLAS (HMat a)` (BLAS/HMat.hs:103-231):
--
--  * `ElemB CuMat` is a small SYMBOLIC scalar (`Sc`).  `liftB` receives a Haskell closure @Vec n a -> a@; a GPU cannot
--    call it per element, so the closure is applied ONCE to symbolic variables and the resulting expression is
--    shipped as postfix bytecode to `tops_lift` (include/tops_b200.h, TOPS_OP_*).  `ad`'s `diff`/`grad`
--    (TOp.hs:212,246) work unchanged because they only need `Floating`.
--  * Scalars that the class returns by value (`dot`, `traceB`, `sumB`, `indexB`) are `Lit` leaves read back with
--    `tops_index` — the reference's observation points, the only calls that synchronise.
module TensorOps.BLAS.Cuda
  ( CuMat
  , Sc(..)
  , withCuda
  -- * shared with TensorOps.Backend.Cuda (`instance Tensor CuTensor`)
  , Ctx, Buf, Dev(..), theCtx, check, newOut, pureOut, withDev, scalarOf, compile, shapeOf, download, unLit
  ) where

import           Data.IORef
import           Data.Int
import           Data.Kind                 (Type)
import           Data.Singletons
import           Data.Type.Vector          (Vec, VecT (..), I (..))
import           Foreign
import           Foreign.C.String
import           Foreign.C.Types
import           System.IO.Unsafe          (unsafePerformIO)
import           Unsafe.Coerce             (unsafeCoerce)
import           TensorOps.BLAS
import           TensorOps.NatKind
import qualified Data.Finite               as DF

-- ---------------------------------------------------------------------------------------------------------------
-- raw bindings (one per entry point of include/tops_b200.h that the class needs)

data Ctx
data Buf

foreign import ccall unsafe "tops_init"        c_init        :: CInt -> Ptr (Ptr Ctx) -> IO CInt
foreign import ccall unsafe "tops_last_error"  c_last_error  :: Ptr Ctx -> IO CString
foreign import ccall unsafe "tops_buf_alloc"   c_buf_alloc   :: Ptr Ctx -> CInt -> CInt -> Ptr Int64 -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "&tops_buf_release" p_buf_release :: FunPtr (Ptr Buf -> IO ())
foreign import ccall unsafe "tops_upload"      c_upload      :: Ptr Ctx -> Ptr Buf -> Ptr CFloat -> CSize -> IO CInt
foreign import ccall safe   "tops_download"    c_download    :: Ptr Ctx -> Ptr Buf -> Ptr CFloat -> CSize -> IO CInt
foreign import ccall unsafe "tops_buf_numel"   c_buf_numel   :: Ptr Buf -> IO Int64
foreign import ccall unsafe "tops_buf_rank"    c_buf_rank    :: Ptr Buf -> IO CInt
foreign import ccall unsafe "tops_buf_dims"    c_buf_dims    :: Ptr Buf -> Ptr Int64 -> IO CInt
foreign import ccall unsafe "tops_fill"        c_fill        :: Ptr Ctx -> Ptr Buf -> CDouble -> IO CInt
foreign import ccall unsafe "tops_axpy"        c_axpy        :: Ptr Ctx -> CDouble -> Ptr Buf -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_dot"         c_dot         :: Ptr Ctx -> Ptr Buf -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_ger"         c_ger         :: Ptr Ctx -> Ptr Buf -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_gemv"        c_gemv        :: Ptr Ctx -> CDouble -> Ptr Buf -> Ptr Buf -> CDouble -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_gemm"        c_gemm        :: Ptr Ctx -> CDouble -> Ptr Buf -> Ptr Buf -> CDouble -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_scale"       c_scale       :: Ptr Ctx -> CDouble -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_add"         c_add         :: Ptr Ctx -> Ptr Buf -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall safe   "tops_index"       c_index       :: Ptr Ctx -> Ptr Buf -> Ptr Int64 -> Ptr CDouble -> IO CInt
foreign import ccall unsafe "tops_index_row"   c_index_row   :: Ptr Ctx -> Ptr Buf -> Int64 -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_transp"      c_transp      :: Ptr Ctx -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_eye"         c_eye         :: Ptr Ctx -> Int64 -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_trace"       c_trace       :: Ptr Ctx -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_diag"        c_diag        :: Ptr Ctx -> CInt -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_get_diag"    c_get_diag    :: Ptr Ctx -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_sum"         c_sum         :: Ptr Ctx -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_lift"        c_lift        :: Ptr Ctx -> Ptr Int32 -> CInt -> Ptr CFloat -> CInt -> CInt
                                                              -> Ptr (Ptr Buf) -> CInt -> Ptr Int64 -> Ptr (Ptr Buf) -> IO CInt

-- ---------------------------------------------------------------------------------------------------------------
-- context: one per process, created on first use (the class methods are pure, so there is nowhere to thread it)

{-# NOINLINE theCtx #-}
theCtx :: Ptr Ctx
theCtx = unsafePerformIO $ alloca $ \pp -> do
    rc <- c_init 0 pp
    if rc /= 0 then error ("tops_init failed with status " ++ show rc ++ " (no sm_100 device? there is no CPU fallback)")
               else peek pp

-- | Bracket for applications that want the failure at start-up rather than at first use.
withCuda :: IO a -> IO a
withCuda act = theCtx `seq` act

check :: String -> CInt -> IO ()
check what rc
  | rc == 0   = return ()
  | otherwise = do msg <- peekCString =<< c_last_error theCtx
                   error (what ++ ": " ++ msg)

-- | A device tensor: a ref-counted `tops_buf`; the finalizer is `tops_buf_release`, whose free is stream-ordered
--   (it may run while kernels that read the buffer are still in flight).
newtype Dev = Dev (ForeignPtr Buf)

newOut :: String -> (Ptr (Ptr Buf) -> IO CInt) -> IO Dev
newOut what call = alloca $ \pp -> do
    poke pp nullPtr
    check what =<< call pp
    Dev <$> (newForeignPtr p_buf_release =<< peek pp)

withDev :: Dev -> (Ptr Buf -> IO a) -> IO a
withDev (Dev fp) = withForeignPtr fp

pureOut :: String -> (Ptr (Ptr Buf) -> IO CInt) -> Dev
pureOut what call = unsafePerformIO (newOut what call)
{-# NOINLINE pureOut #-}

-- | read a rank-0 result back (synchronises: this is an observation point like `atIndex` in HMat.hs:163-169)
scalarOf :: Dev -> Float
scalarOf d = unsafePerformIO $ withDev d $ \b -> alloca $ \pv -> do
    check "tops_index" =<< c_index theCtx b nullPtr pv
    realToFrac <$> peek pv

-- ---------------------------------------------------------------------------------------------------------------
-- the symbolic element type

data Sc = Lit !Float | Var !Int | Un !Int32 Sc | Bin !Int32 Sc Sc

opAdd, opSub, opMul, opDiv, opNeg, opExp, opLog, opRecip, opSqrt, opTanh, opAbs, opSignum, opPow, opSin, opCos :: Int32
opAdd = 2; opSub = 3; opMul = 4; opDiv = 5; opNeg = 6; opExp = 7; opLog = 8; opRecip = 9; opSqrt = 10; opTanh = 11
opAbs = 12; opSignum = 13; opPow = 16; opSin = 18; opCos = 19

instance Num Sc where
    Lit a + Lit b = Lit (a + b); a + b = Bin opAdd a b
    Lit a - Lit b = Lit (a - b); a - b = Bin opSub a b
    Lit a * Lit b = Lit (a * b); a * b = Bin opMul a b
    negate (Lit a) = Lit (negate a); negate a = Un opNeg a
    abs    (Lit a) = Lit (abs a);    abs a    = Un opAbs a
    signum (Lit a) = Lit (signum a); signum a = Un opSignum a
    fromInteger = Lit . fromInteger
instance Fractional Sc where
    Lit a / Lit b = Lit (a / b); a / b = Bin opDiv a b
    recip (Lit a) = Lit (recip a); recip a = Un opRecip a
    fromRational = Lit . fromRational
instance Floating Sc where
    pi = Lit pi
    exp  (Lit a) = Lit (exp a);  exp a  = Un opExp a
    log  (Lit a) = Lit (log a);  log a  = Un opLog a
    sqrt (Lit a) = Lit (sqrt a); sqrt a = Un opSqrt a
    tanh (Lit a) = Lit (tanh a); tanh a = Un opTanh a
    sin  (Lit a) = Lit (sin a);  sin a  = Un opSin a
    cos  (Lit a) = Lit (cos a);  cos a  = Un opCos a
    Lit a ** Lit b = Lit (a ** b); a ** b = Bin opPow a b
    -- the remaining methods are definable from the above (asin .. atanh): left as `error` until a TOp needs them
    asin = unsup "asin"; acos = unsup "acos"; atan = unsup "atan"; sinh = unsup "sinh"; cosh = unsup "cosh"
    asinh = unsup "asinh"; acosh = unsup "acosh"; atanh = unsup "atanh"

unsup :: String -> a
unsup f = error ("TensorOps.BLAS.Cuda: " ++ f ++ " has no TOPS_OP_* opcode yet")

-- The reference asks for MORE than `Floating`: `instance Tensor (BTensor v b)` needs `RealFloat (ElemB b)` (BTensor.hs:775-785),
-- `TOp`'s fields are `forall t. (Tensor t, RealFloat (ElemT t)) => ...` (Types.hs:122-125) and `VFunc` closures are
-- `forall a. RealFloat a` (Types.hs:114-117).  `RealFloat` drags in Eq / Ord / Real / RealFrac, which a symbolic expression can
-- only answer for literals: every method below is total on `Lit` and an `error` on a symbolic argument — a closure that BRANCHES
-- on its tensor argument (compares, rounds, decodes it) cannot be reified into one elementwise program and is reported as such.
-- None of the reference's own TOps do (logistic, softmax, the losses, `ad`'s `diff`/`grad` use only Num/Fractional/Floating).
symbolic :: String -> a
symbolic f = error ("TensorOps.BLAS.Cuda: " ++ f ++ " on a symbolic element — the lifted closure inspects its argument instead of computing with it")

instance Eq Sc where
    Lit a == Lit b = a == b
    _     == _     = symbolic "(==)"
instance Ord Sc where
    compare (Lit a) (Lit b) = compare a b
    compare _       _       = symbolic "compare"
    -- `max` / `min` do have opcodes (TOPS_OP_MAX = 14, TOPS_OP_MIN = 15): a ReLU-style `max 0 x` stays liftable
    max (Lit a) (Lit b) = Lit (max a b); max a b = Bin 14 a b
    min (Lit a) (Lit b) = Lit (min a b); min a b = Bin 15 a b
instance Real Sc where
    toRational (Lit a) = toRational a
    toRational _       = symbolic "toRational"
instance RealFrac Sc where
    properFraction (Lit a) = let (n, f) = properFraction a in (n, Lit f)
    properFraction _       = symbolic "properFraction"
instance RealFloat Sc where
    floatRadix     _ = floatRadix     (0 :: Float)
    floatDigits    _ = floatDigits    (0 :: Float)
    floatRange     _ = floatRange     (0 :: Float)
    decodeFloat (Lit a) = decodeFloat a
    decodeFloat _       = symbolic "decodeFloat"
    encodeFloat m e     = Lit (encodeFloat m e)
    isNaN          = litOnly "isNaN" isNaN
    isInfinite     = litOnly "isInfinite" isInfinite
    isDenormalized = litOnly "isDenormalized" isDenormalized
    isNegativeZero = litOnly "isNegativeZero" isNegativeZero
    isIEEE         _ = True
    atan2 (Lit a) (Lit b) = Lit (atan2 a b)
    atan2 _       _       = unsup "atan2"

litOnly :: String -> (Float -> Bool) -> Sc -> Bool
litOnly _ f (Lit a) = f a
litOnly n _ _       = symbolic n

-- | postfix program + constant pool for `tops_lift`
compile :: Sc -> ([Int32], [Float])
compile e0 = let (code, cs) = go e0 ([], []) in (reverse code, reverse cs)
  where
    ins op arg = (op `shiftL` 16) .|. arg
    go (Var i)     (c, k) = (ins 0 (fromIntegral i) : c, k)
    go (Lit v)     (c, k) = (ins 1 (fromIntegral (length k)) : c, v : k)
    go (Un op a)   s      = let (c, k) = go a s in (ins op 0 : c, k)
    go (Bin op a b) s     = let (c, k) = go b (go a s) in (ins op 0 : c, k)

-- ---------------------------------------------------------------------------------------------------------------
-- the instance

-- | `CuMat s`: a device vector ('BV n) or row-major matrix ('BM n m).  The shape is carried at run time by the buffer.
newtype CuMat (s :: BShape Nat) = CuMat Dev

dimsOf :: Sing (s :: BShape Nat) -> [Int64]
dimsOf = \case
    SBV n   -> [fromIntegral (fromSing n)]
    SBM n m -> [fromIntegral (fromSing n), fromIntegral (fromSing m)]

vecToList :: Vec n a -> [a]
vecToList = \case { ØV -> []; I x :* xs -> x : vecToList xs }

symbolicArgs :: Vec n b -> Vec n Sc
symbolicArgs = go 0
  where
    go :: Int -> Vec m b -> Vec m Sc
    go _ ØV        = ØV
    go i (_ :* xs) = I (Var i) :* go (i + 1) xs

instance BLAS CuMat where
    type ElemB CuMat = Sc

    -- liftB (BLAS.hs:92-96; reference body HMat.hs:108-133)
    liftB s f xs =
        let (code, consts) = compile (f (symbolicArgs xs))
            ins            = [ d | CuMat d <- vecToList xs ]
            ds             = dimsOf s
        in  CuMat $ pureOut "tops_lift" $ \out ->
              withArrayLen code $ \nc pc ->
              withArrayLen (map realToFrac consts) $ \nk pk ->
              withMany withDev ins $ \bs ->
              withArrayLen bs $ \ni pin ->
              withArrayLen ds $ \r pd ->
                c_lift theCtx pc (fromIntegral nc) pk (fromIntegral nk) (fromIntegral ni) pin (fromIntegral r) pd out

    -- axpy (BLAS.hs:97-101; HMat.hs:135-139)
    axpy (Lit a) (CuMat x) my = CuMat $ pureOut "tops_axpy" $ \out ->
        withDev x $ \px -> maybe ($ nullPtr) (\(CuMat y) -> withDev y) my $ \py ->
          c_axpy theCtx (realToFrac a) px py out
    axpy _ _ _ = error "axpy: symbolic alpha"

    dot (CuMat x) (CuMat y) = Lit . scalarOf $ pureOut "tops_dot" $ \out ->
        withDev x $ \px -> withDev y $ \py -> c_dot theCtx px py out

    ger (CuMat x) (CuMat y) = CuMat $ pureOut "tops_ger" $ \out ->
        withDev x $ \px -> withDev y $ \py -> c_ger theCtx px py out

    -- gemv / gemm (BLAS.hs:111-123; HMat.hs:147-160): alpha/beta are fused in the kernel, not three passes
    gemv (Lit a) (CuMat m) (CuMat x) mby = CuMat $ pureOut "tops_gemv" $ \out ->
        withDev m $ \pm -> withDev x $ \px ->
          case mby of
            Nothing               -> c_gemv theCtx (realToFrac a) pm px 0 nullPtr out
            Just (Lit b, CuMat y) -> withDev y $ \py -> c_gemv theCtx (realToFrac a) pm px (realToFrac b) py out
            _                     -> error "gemv: symbolic beta"
    gemv _ _ _ _ = error "gemv: symbolic alpha"

    gemm (Lit a) (CuMat x) (CuMat y) mbc = CuMat $ pureOut "tops_gemm" $ \out ->
        withDev x $ \px -> withDev y $ \py ->
          case mbc of
            Nothing               -> c_gemm theCtx (realToFrac a) px py 0 nullPtr out
            Just (Lit b, CuMat c) -> withDev c $ \pc -> c_gemm theCtx (realToFrac a) px py (realToFrac b) pc out
            _                     -> error "gemm: symbolic beta"
    gemm _ _ _ _ = error "gemm: symbolic alpha"

    scaleB (Lit a) (CuMat x) = CuMat $ pureOut "tops_scale" $ \out -> withDev x $ \px -> c_scale theCtx (realToFrac a) px out
    scaleB _ _ = error "scaleB: symbolic alpha"
    addB (CuMat x) (CuMat y) = CuMat $ pureOut "tops_add" $ \out -> withDev x $ \px -> withDev y $ \py -> c_add theCtx px py out

    indexB ix (CuMat x) = Lit $ unsafePerformIO $ withDev x $ \px ->
        withArray (case ix of { PBV i -> [fin i]; PBM i j -> [fin i, fin j] }) $ \pi' -> alloca $ \pv -> do
          check "tops_index" =<< c_index theCtx px pi' pv
          realToFrac <$> peek pv
      where fin = fromInteger . DF.getFinite

    indexRowB i (CuMat x) = CuMat $ pureOut "tops_index_row" $ \out ->
        withDev x $ \px -> c_index_row theCtx px (fromInteger (DF.getFinite i)) out

    -- O(1): flips the storage-order flag of a view, like hmatrix `tr` (HMat.hs:175)
    transpB (CuMat x) = CuMat $ pureOut "tops_transp" $ \out -> withDev x $ \px -> c_transp theCtx px out

    eye n = CuMat $ pureOut "tops_eye" $ \out -> c_eye theCtx (fromIntegral (fromSing n)) out
    traceB (CuMat x) = Lit . scalarOf $ pureOut "tops_trace" $ \out -> withDev x $ \px -> c_trace theCtx px out
    diagB (CuMat x) = CuMat $ pureOut "tops_diag" $ \out -> withDev x $ \px -> c_diag theCtx 2 px out
    getDiagB (CuMat x) = CuMat $ pureOut "tops_get_diag" $ \out -> withDev x $ \px -> c_get_diag theCtx px out
    sumB (CuMat x) = Lit . scalarOf $ pureOut "tops_sum" $ \out -> withDev x $ \px -> c_sum theCtx px out

    -- Element-at-a-time generators / traversals (BLAS.hs:140-157; HMat.hs:176-216) are host<->device by nature:
    -- build the host array once, one upload (bgenA) / one download + one upload (iElemsB, iRowsB, bgenRowsA).
    bgenA s f = upload (dimsOf s) <$> traverse (fmap unLit . f) (indices s)

    bgenRowsA :: forall f n m. (Applicative f, SingI n) => (DF.Finite n -> f (CuMat ('BV m))) -> f (CuMat ('BM n m))
    bgenRowsA f = stackRows <$> traverse (f . DF.finite) [0 .. fromSing (sing :: Sing n) - 1]

    iRowsB f m@(CuMat d) = stackRows <$> traverse (\i -> f (DF.finite i) (indexRowB (DF.finite i) m)) [0 .. nRows - 1]
      where nRows = fromIntegral (head (shapeOf d))

    iElemsB f x@(CuMat d) = upload ds <$> traverse (\(ix, e) -> unLit <$> f ix (Lit e)) (zip (indicesOfDims ds) (download d))
      where ds = shapeOf d

-- | rows (device vectors) -> one device matrix: each row is downloaded and the matrix uploaded once (setup-time path only)
stackRows :: [CuMat ('BV m)] -> CuMat ('BM n m)
stackRows rows = upload [fromIntegral (length rows), fromIntegral (length (head hostRows))] (concat hostRows)
  where hostRows = [ download d | CuMat d <- rows ]

-- | logical dims of a device tensor (tops_buf_rank / tops_buf_dims)
shapeOf :: Dev -> [Int64]
shapeOf d = unsafePerformIO $ withDev d $ \b -> do
    r <- c_buf_rank b
    allocaArray (fromIntegral r) $ \pd -> c_buf_dims b pd >> peekArray (fromIntegral r) pd

-- | the whole tensor as a row-major host list (tops_download synchronises)
download :: Dev -> [Float]
download d = unsafePerformIO $ withDev d $ \b -> do
    n <- fromIntegral <$> c_buf_numel b
    allocaArray n $ \p -> do
      check "tops_download" =<< c_download theCtx b p (fromIntegral (4 * n))
      map realToFrac <$> peekArray n p

indicesOfDims :: [Int64] -> [BShapeP DF.Finite s]
indicesOfDims = \case
    [n]    -> unsafeCoerce [ PBV (DF.finite (fromIntegral i)) | i <- [0 .. n - 1] ]
    [n, m] -> unsafeCoerce [ PBM (DF.finite (fromIntegral i)) (DF.finite (fromIntegral j)) | i <- [0 .. n - 1], j <- [0 .. m - 1] ]
    _      -> error "indicesOfDims: BLAS shapes are vectors or matrices"

unLit :: Sc -> Float
unLit (Lit v) = v
unLit _       = error "generator produced a symbolic scalar"

indices :: Sing (s :: BShape Nat) -> [BShapeP DF.Finite s]
indices = \case
    SBV n   -> [ PBV (DF.finite i) | i <- [0 .. fromSing n - 1] ]
    SBM n m -> [ PBM (DF.finite i) (DF.finite j) | i <- [0 .. fromSing n - 1], j <- [0 .. fromSing m - 1] ]

upload :: [Int64] -> [Float] -> CuMat s
upload ds xs = CuMat $ unsafePerformIO $ do
    d <- newOut "tops_buf_alloc" $ \out -> withArrayLen ds $ \r pd -> c_buf_alloc theCtx 0 (fromIntegral r) pd out
    withDev d $ \b -> withArrayLen (map realToFrac xs) $ \n p ->
      check "tops_upload" =<< c_upload theCtx b p (fromIntegral (4 * n))
    return d
