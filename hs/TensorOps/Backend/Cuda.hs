{-# LANGUAGE DataKinds                #-}
{-# LANGUAGE FlexibleContexts         #-}
{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE GADTs                    #-}
{-# LANGUAGE InstanceSigs             #-}
{-# LANGUAGE KindSignatures           #-}
{-# LANGUAGE LambdaCase               #-}
{-# LANGUAGE PolyKinds                #-}
{-# LANGUAGE RankNTypes               #-}
{-# LANGUAGE ScopedTypeVariables      #-}
{-# LANGUAGE TypeApplications         #-}
{-# LANGUAGE TypeFamilies             #-}
{-# LANGUAGE TypeOperators            #-}

-- | TensorOps.Backend.Cuda — `instance Tensor CuTensor`: the reference's `Tensor` dictionary (src/TensorOps/Types.hs:52-109) bound
-- method by method to the flat-storage entry points of libtops_b200.so (include/tops_b200.h).
--
-- Drop into the reference tree as @src/TensorOps/Backend/Cuda.hs@ next to BLAS/Cuda.hs; select it by type application where the
-- apps select hmatrix today: @Proxy \@CuTensor@ instead of @Proxy \@(BTensorV (HMat Double))@ (app/Dots.hs:145, app/MNIST.hs:154).
--
-- What it changes against `BTensor v (HMat a)`:
--
--  * a tensor of ANY rank is one contiguous device buffer (row-major, first index outermost) — no nested boxed vectors
--    (Backend/BTensor.hs:58-63), so `gmul` on any ranks is ONE call (tops_gmul: a single tensor-core GEMM on M = prod ms,
--    K = prod os, N = prod ns) instead of the rank dispatch + per-matrix BLAS calls + `naiveGMul` of BTensor.hs:592-716;
--  * rank-0 results (`dot`, losses, `sumRows` of a vector) STAY ON THE DEVICE as rank-0 buffers; only `(!)` reads a value back —
--    a per-sample loss no longer round-trips to the host between the forward and the reverse sweep;
--  * `liftT` closures are applied once to symbolic `Sc` variables and shipped as bytecode (see BLAS/Cuda.hs);
--  * `sumT` of matrices is an n-ary add, not the `gemm 1 xs (eye m) (Just (1, ys))` of BTensor.hs:113.
--
-- STATUS: written against include/tops_b200.h and the class definition; NOT COMPILED — no GHC in the build image (SURVEY.md
-- fact 5).  The C ABI it calls is what tests/ exercise (through ctypes, with the same argument conventions).
module TensorOps.Backend.Cuda
  ( CuTensor
  ) where

import           Control.Monad.Primitive
import           Data.Int
import           Data.Kind                 (Type)
import           Data.Singletons
import           Data.Singletons.Prelude   (Sing (..))
import           Data.Type.Index
import           Data.Type.Length
import           Data.Type.Sing             (takeSing)
import           Data.Type.Product
import           Data.Type.Uniform
import           Data.Type.Vector          (Vec, VecT (..), I (..))
import           Foreign
import           Foreign.C.Types
import           GHC.TypeLits              (Nat)
import           Statistics.Distribution   (ContGen (..))
import           System.IO.Unsafe          (unsafePerformIO)
import           TensorOps.BLAS.Cuda
import           TensorOps.Types
import           Type.Family.List
import qualified Data.Finite               as DF

foreign import ccall unsafe "tops_buf_alloc"     c_buf_alloc     :: Ptr Ctx -> CInt -> CInt -> Ptr Int64 -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_upload"        c_upload        :: Ptr Ctx -> Ptr Buf -> Ptr CFloat -> CSize -> IO CInt
foreign import ccall unsafe "tops_lift"          c_lift          :: Ptr Ctx -> Ptr Int32 -> CInt -> Ptr CFloat -> CInt -> CInt
                                                                  -> Ptr (Ptr Buf) -> CInt -> Ptr Int64 -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_gmul"          c_gmul          :: Ptr Ctx -> CInt -> CInt -> CInt -> Ptr Buf -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_sum_t"         c_sum_t         :: Ptr Ctx -> CInt -> Ptr (Ptr Buf) -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_scale"         c_scale         :: Ptr Ctx -> CDouble -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_transp"        c_transp        :: Ptr Ctx -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_sum_rows"      c_sum_rows      :: Ptr Ctx -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_diag"          c_diag          :: Ptr Ctx -> CInt -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_get_diag"      c_get_diag      :: Ptr Ctx -> Ptr Buf -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_buf_view"      c_buf_view      :: Ptr Ctx -> Ptr Buf -> Int64 -> CInt -> Ptr Int64 -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall unsafe "tops_copy"          c_copy          :: Ptr Ctx -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe   "tops_index"         c_index         :: Ptr Ctx -> Ptr Buf -> Ptr Int64 -> Ptr CDouble -> IO CInt

-- | A device tensor of shape @ns@; the shape lives in the type (and in the buffer's own header for the C side).
newtype CuTensor (ns :: [Nat]) = CuT Dev

dimsS :: Sing (ns :: [Nat]) -> [Int64]
dimsS = map fromIntegral . fromSing

lenI :: Length (as :: [k]) -> Int
lenI = \case { LZ -> 0; LS l -> 1 + lenI l }

ixs :: Prod DF.Finite (ns :: [Nat]) -> [Int64]
ixs = \case { Ø -> []; i :< is -> fromInteger (DF.getFinite i) : ixs is }

vecToList :: Vec n a -> [a]
vecToList = \case { ØV -> []; I x :* xs -> x : vecToList xs }

symbolicArgs :: Vec n b -> Vec n Sc
symbolicArgs = go 0
  where
    go :: Int -> Vec m b -> Vec m Sc
    go _ ØV        = ØV
    go i (_ :* xs) = I (Var i) :* go (i + 1) xs

out1 :: String -> (Ptr (Ptr Buf) -> IO CInt) -> CuTensor ns
out1 what call = CuT (pureOut what call)

-- | a flat sub-tensor view: `offset` elements into the storage, `dims` its shape (tops_buf_view keeps the parent alive)
viewAt :: Dev -> Int64 -> [Int64] -> Dev
viewAt d off dims = pureOut "tops_buf_view" $ \out -> withDev d $ \p ->
    withArrayLen dims $ \r pd -> c_buf_view theCtx p off (fromIntegral r) pd out

uploadT :: [Int64] -> [Float] -> CuTensor ns
uploadT ds xs = CuT $ unsafePerformIO $ do
    d <- newOut "tops_buf_alloc" $ \out -> withArrayLen ds $ \r pd -> c_buf_alloc theCtx 0 (fromIntegral r) pd out
    withDev d $ \b -> withArrayLen (map realToFrac xs) $ \n p ->
      check "tops_upload" =<< c_upload theCtx b p (fromIntegral (4 * n))
    return d

-- | all index tuples of a shape in row-major order (first index outermost) — the storage order
allIndices :: Sing (ns :: [Nat]) -> [Prod DF.Finite ns]
allIndices = \case
    SNil         -> [Ø]
    n `SCons` ns -> [ DF.finite i :< is | i <- [0 .. fromIntegral (fromSing n) - 1], is <- allIndices ns ]

instance Tensor CuTensor where
    type ElemT CuTensor = Sc

    -- Types.hs:56-59 / TOp.hs:42-54: the closure is reified once; tops_lift interprets (or dispatches to a specialised kernel)
    liftT :: forall o n. SingI o => (Vec n Sc -> Sc) -> Vec n (CuTensor o) -> CuTensor o
    liftT f xs =
        let (code, consts) = compile (f (symbolicArgs xs))
            ins            = [ d | CuT d <- vecToList xs ]
            ds             = dimsS (sing :: Sing o)
        in  out1 "tops_lift" $ \out ->
              withArrayLen code $ \nc pc ->
              withArrayLen (map realToFrac consts) $ \nk pk ->
              withMany withDev ins $ \bs ->
              withArrayLen bs $ \ni pin ->
              withArrayLen ds $ \r pd ->
                c_lift theCtx pc (fromIntegral nc) pk (fromIntegral nk) (fromIntegral ni) pin (fromIntegral r) pd out

    -- Types.hs:60-66: y's contraction axes are REVERSED; tops_gmul has exactly this convention
    gmul lM lO lN (CuT x) (CuT y) = out1 "tops_gmul" $ \out ->
        withDev x $ \px -> withDev y $ \py ->
          c_gmul theCtx (fromIntegral (lenI lM)) (fromIntegral (lenI lO)) (fromIntegral (lenI lN)) px py out

    sumT [] = error "sumT []: the class gives no shape to build a zero tensor from at run time without SingI evidence in scope"
    sumT ts = out1 "tops_sum_t" $ \out ->
        withMany withDev [ d | CuT d <- ts ] $ \bs -> withArrayLen bs $ \n pb -> c_sum_t theCtx (fromIntegral n) pb out

    scaleT (Lit a) (CuT x) = out1 "tops_scale" $ \out -> withDev x $ \px -> c_scale theCtx (realToFrac a) px out
    scaleT _       _       = error "scaleT: symbolic factor"

    -- Types.hs:71-73: full axis reversal (an O(1) view for rank <= 2, one permutation kernel above)
    transp (CuT x) = out1 "tops_transp" $ \out -> withDev x $ \px -> c_transp theCtx px out

    -- Types.hs:77-81: f is applied to each sub-tensor VIEW (no host round trip of the data); the results are gathered with
    -- device-to-device copies into one buffer
    mapRows :: forall ns ms. SingI (ns ++ ms) => Length ns -> (CuTensor ms -> CuTensor ms) -> CuTensor (ns ++ ms) -> CuTensor (ns ++ ms)
    mapRows lN f (CuT x) = CuT $ unsafePerformIO $ do
        let ds        = dimsS (sing :: Sing (ns ++ ms))
            (hd, tl)  = splitAt (lenI lN) ds
            rows      = product hd
            rowElems  = product tl
        out <- newOut "tops_buf_alloc" $ \o -> withArrayLen ds $ \r pd -> c_buf_alloc theCtx 0 (fromIntegral r) pd o
        mapM_ (\i -> do let CuT y = f (CuT (viewAt x (i * rowElems) tl))
                        withDev (viewAt out (i * rowElems) tl) $ \pd' -> withDev y $ \py ->
                          check "tops_copy" =<< c_copy theCtx pd' py)
              [0 .. rows - 1]
        return out

    sumRows (CuT x) = out1 "tops_sum_rows" $ \out -> withDev x $ \px -> c_sum_rows theCtx px out

    diag :: forall n ns. SingI (n ': ns) => Uniform n ns -> CuTensor '[n] -> CuTensor (n ': ns)
    diag _ (CuT v) = out1 "tops_diag" $ \out -> withDev v $ \pv ->
        c_diag theCtx (fromIntegral (length (dimsS (sing :: Sing (n ': ns))))) pv out

    getDiag _ (CuT x) = out1 "tops_get_diag" $ \out -> withDev x $ \px -> c_get_diag theCtx px out

    -- BTensor.hs:838-842 draws every element through `generateA`; here too (setup-time path), with ONE upload.
    -- (tops_rand_normal / tops_rand_uniform generate on the device when the distribution is known to be one of those.)
    genRand d g = generateA (\_ -> Lit . realToFrac <$> genContVar d g)

    generateA :: forall f ns. (Applicative f, SingI ns) => (Prod DF.Finite ns -> f Sc) -> f (CuTensor ns)
    generateA f = uploadT (dimsS s) . map unLit <$> traverse f (allIndices s)
      where s = sing :: Sing ns

    -- Types.hs:99-106: traverse the leading `ms` axes; each visited sub-tensor is a device view, each result a device tensor
    ixRows :: forall f ms os ns. (Applicative f, SingI (ms ++ os))
           => Length ms -> Length os -> (Prod DF.Finite ms -> CuTensor ns -> f (CuTensor os)) -> CuTensor (ms ++ ns) -> f (CuTensor (ms ++ os))
    ixRows lM lO f (CuT x) = gather <$> traverse (\(k, ix) -> f ix (CuT (viewAt x (k * inElems) inDims))) (zip [0 ..] idxs)
      where
        outAll          = dimsS (sing :: Sing (ms ++ os))
        oDims           = drop (lenI lM) outAll
        xDims           = shapeOf x
        inDims          = drop (lenI lM) xDims
        inElems         = product inDims
        outElems        = product oDims
        idxs            = allIndices (takeSing lM lO (sing :: Sing (ms ++ os)))   -- as BTensor.hs does for its own ixRows
        gather rs = CuT $ unsafePerformIO $ do
            out <- newOut "tops_buf_alloc" $ \o -> withArrayLen outAll $ \r pd -> c_buf_alloc theCtx 0 (fromIntegral r) pd o
            mapM_ (\(k, CuT y) -> withDev (viewAt out (k * outElems) oDims) $ \pd' -> withDev y $ \py ->
                                    check "tops_copy" =<< c_copy theCtx pd' py)
                  (zip [0 ..] rs)
            return out

    -- the one observation point: a single element crosses PCIe (tops_index synchronises)
    CuT x ! ix = Lit $ unsafePerformIO $ withDev x $ \px -> withArray (ixs ix) $ \pi' -> alloca $ \pv -> do
        check "tops_index" =<< c_index theCtx px pi' pv
        realToFrac <$> peek pv
