;;;; mfit-cuda.lisp -- SBCL host of the B200 backend: the reference's walker API over libmfit_cuda.so (CFFI).
;;;;
;;;; Drop-in for the hot path of afranson/Lisp-MCMC (mcmc-fitting.lisp): walker-create (:1132-1163),
;;;; walker-adaptive-steps(-full) (:862-947), walker-many-steps (:849-853), walker-sample-region (:949-969),
;;;; walker-force-take-step (:1124-1129), the trace-reading modes of walker-get (:487-543) and walker-save
;;;; (:971-1001, dead code in the reference).  Same function names and keyword arguments; the three plugin
;;;; slots take *descriptors* instead of closures, because nothing on the GPU path calls back into Lisp:
;;;;
;;;;   :function       (gpu-lambda (x &key a b &allow-other-keys) (+ b (* a x)))   ; any reference-style model,
;;;;                                                                               ; lowered to CUDA C -> NVRTC -> sm_100a
;;;;                   (gpu-builtin :gauss-lin '(:a :mu :s :b0 :b1))               ; hand-written FP64 family
;;;;   :log-liklihood  :log-liklihood-normal | :log-liklihood-normal-cutoff | :log-poisson
;;;;                   | (create-log-liklihood-function (gpu-lambda (y model error) (log-poisson model y)))  ; :402-417, any
;;;;                                                        ; per-point lambda; ERROR is the whole stddev list as in :416
;;;;   :log-prior      :log-prior-flat | ((:key lo hi) ...)  ; = (prior-bounds-let (...) bounds-total)
;;;;                   | (gpu-lambda (params data) (prior-bounds-let (...) (+ bounds-total ...)))
;;;;
;;;; Several GPUs: (walker-create ... :devices '(0 1 2 3) :shard :chains) shards the W chains over the listed GPUs of this
;;;; box, :shard :datasets puts data set d of a global fit on GPU d mod n; both inside the library (mfit_multi_*, one host
;;;; thread per GPU), bit-identical to one GPU.  The reference has one worker (`mapc`, nv-specific.lisp:58-66).
;;;;
;;;; A gpu-walker holds W independent chains (a "walker set", nv-specific.lisp:58-66) that share
;;;; :function / :data / :log-liklihood / :log-prior; W = 1 is exactly one reference walker.
;;;;
;;;; NOT EXECUTED IN THE BUILD IMAGE (no Common Lisp there): the Python mirror lisp_mcmc_b200/walker.py sits on
;;;; the same C ABI and is what the parity tests drive; tests/test_abi_and_host.py checks that every defcfun
;;;; below names a symbol of include/mfit_cuda.h with the declared number of arguments.
;;;;
;;;; SBCL notes: float traps are masked around every foreign call (CUDA / NVRTC code may raise :inexact or
;;;; :invalid); Lisp vectors are copied into foreign memory for the duration of a call (the library copies
;;;; again, include/mfit_cuda.h "ownership"), so no pinning and no foreign->Lisp re-entry is needed.

(defpackage :mfit-cuda
  (:use :cl)
  (:export #:gpu-lambda #:gpu-builtin #:gpu-walker #:create-log-liklihood-function #:walker-create #:walker-adaptive-steps
           #:walker-adaptive-steps-full #:walker-many-steps #:walker-sample-region #:walker-force-take-step
           #:walker-get #:walker-get-data-and-fit #:walker-save #:walker-load #:walker-modify #:walker-set-get #:walker-free #:mcmc-fit #:mfit-walker-estop #:lower-to-cuda #:walker-set-create))
(in-package :mfit-cuda)

(cffi:define-foreign-library libmfit-cuda (t (:default "libmfit_cuda")))
(cffi:use-foreign-library libmfit-cuda)

;;; ---------------------------------------------------------------------------------- C ABI (include/mfit_cuda.h)
(cffi:defcstruct adapt-cfg                ; mfit_adapt_cfg = keyword arguments of walker-adaptive-steps-full :862
  (n :int64) (temperature :double) (auto-mode :int32) (sampling-optimization :int32)
  (max-walker-length :int64) (greedy :int32) (reserved :int32))

(cffi:defcstruct chain-stats              ; mfit_chain_stats = scalar slots of struct walker :467-479
  (length :int64) (age :int64) (steps-this-run :int64) (last-prob :double) (best-prob :double)
  (temperature :double) (accept-200 :double) (shutting-down :int32) (l-updates :int32))

(cffi:defcfun ("mfit_abi_version" %abi-version) :int)
(cffi:defcfun ("mfit_ctx_create" %ctx-create) :int (device :int) (stream :pointer) (out :pointer))
(cffi:defcfun ("mfit_ctx_destroy" %ctx-destroy) :void (ctx :pointer))
(cffi:defcfun ("mfit_last_error" %last-error) :string (ctx :pointer))
(cffi:defcfun ("mfit_problem_define" %problem-define) :int (ctx :pointer) (n-datasets :int) (n-params :int))
(cffi:defcfun ("mfit_dataset_upload" %dataset-upload) :int
  (ctx :pointer) (d :int) (x :pointer) (y :pointer) (sigma :pointer) (n :int64))
(cffi:defcfun ("mfit_dataset_upload_xd" %dataset-upload-xd) :int
  (ctx :pointer) (d :int) (x :pointer) (x-dims :int) (y :pointer) (sigma :pointer) (n :int64))
(cffi:defcfun ("mfit_dataset_upload_per_walker" %dataset-upload-per-walker) :int
  (ctx :pointer) (d :int) (x :pointer) (y :pointer) (sigma :pointer) (n :int64) (w :int64) (x-shared :int) (sigma-shared :int))
(cffi:defcfun ("mfit_model_builtin" %model-builtin) :int
  (ctx :pointer) (d :int) (family :int) (degree :int) (param-index :pointer) (n-idx :int))
(cffi:defcfun ("mfit_model_sexpr" %model-sexpr) :int
  (ctx :pointer) (d :int) (lambda-text :string) (key-names :pointer) (n-keys :int))
(cffi:defcfun ("mfit_likelihood_set" %likelihood-set) :int (ctx :pointer) (d :int) (kind :int))
(cffi:defcfun ("mfit_likelihood_sexpr" %likelihood-sexpr) :int (ctx :pointer) (d :int) (lambda-text :string))
(cffi:defcfun ("mfit_prior_flat" %prior-flat) :int (ctx :pointer) (d :int))
(cffi:defcfun ("mfit_prior_bounds" %prior-bounds) :int
  (ctx :pointer) (d :int) (n :int) (idx :pointer) (lo :pointer) (hi :pointer))
(cffi:defcfun ("mfit_prior_sexpr" %prior-sexpr) :int
  (ctx :pointer) (d :int) (lambda-text :string) (key-names :pointer) (n-keys :int))
(cffi:defcfun ("mfit_lower_sexpr" %lower-sexpr) :int
  (lambda-text :string) (key-names :pointer) (n-keys :int) (kind :int)
  (out :pointer) (out-len :int) (err :pointer) (err-len :int))
(cffi:defcfun ("mfit_eval_logp" %eval-logp) :int
  (ctx :pointer) (theta :pointer) (w :int64) (prob :pointer) (ll :pointer) (lp :pointer))
(cffi:defcfun ("mfit_chains_init" %chains-init) :int
  (ctx :pointer) (theta0 :pointer) (w :int64) (seed :uint64) (chain-id-base :int64) (trace-capacity :int64))
(cffi:defcfun ("mfit_run_begin" %run-begin) :int
  (ctx :pointer) (cfg :pointer) (l-matrix :pointer) (z :pointer) (u :pointer) (nstream :int64))
(cffi:defcfun ("mfit_run_advance" %run-advance) :int
  (ctx :pointer) (max-steps :int64) (steps-done :pointer) (finished :pointer))
(cffi:defcfun ("mfit_run_many_steps" %run-many-steps) :int
  (ctx :pointer) (n :int64) (l-matrix :pointer) (z :pointer) (u :pointer) (nstream :int64))
(cffi:defcfun ("mfit_run_many_steps_per_chain" %run-many-steps-per-chain) :int
  (ctx :pointer) (n :int64) (l-matrices :pointer) (z :pointer) (u :pointer) (nstream :int64))
(cffi:defcfun ("mfit_run_sample_region" %run-sample-region) :int
  (ctx :pointer) (initial-scale :double) (z :pointer) (nstream :int64))
(cffi:defcfun ("mfit_force_take_step" %force-take-step) :int (ctx :pointer))
(cffi:defcfun ("mfit_request_stop" %request-stop) :int (ctx :pointer))
(cffi:defcfun ("mfit_trace_read" %trace-read) :int
  (ctx :pointer) (take :int64) (theta :pointer) (prob :pointer) (n-valid :pointer))
(cffi:defcfun ("mfit_stats_read" %stats-read) :int (ctx :pointer) (stats :pointer))
(cffi:defcfun ("mfit_best_read" %best-read) :int (ctx :pointer) (theta :pointer) (prob :pointer))
(cffi:defcfun ("mfit_last_read" %last-read) :int (ctx :pointer) (theta :pointer) (prob :pointer))
(cffi:defcfun ("mfit_l_matrix_read" %l-matrix-read) :int (ctx :pointer) (l :pointer))
(cffi:defcfun ("mfit_keep_walks" %keep-walks) :int (ctx :pointer) (keep :int64))
(cffi:defcfun ("mfit_walker_modify" %walker-modify) :int (ctx :pointer) (mode :int) (number :int64))
(cffi:defcfun ("mfit_chains_restore" %chains-restore) :int
  (ctx :pointer) (rows :int64) (theta :pointer) (prob :pointer) (n-valid :pointer) (age :pointer)
  (best-theta :pointer) (best-prob :pointer))
(cffi:defcfun ("mfit_trace_analytics" %trace-analytics) :int
  (ctx :pointer) (take :int64) (most-likely-theta :pointer) (most-likely-prob :pointer) (median :pointer)
  (acceptance :pointer) (covariance :pointer) (l-matrix :pointer))

(cffi:defcfun ("mfit_fit_envelope" %fit-envelope) :int
  (ctx :pointer) (d :int) (take :int64) (which-solution :int) (x-fit :pointer) (n-fit :int) (y-fit :pointer)
  (y-max :pointer) (y-min :pointer) (params :pointer) (n-distinct :pointer))

;;; several GPUs behind one handle (csrc/multi.cpp): the twins of the calls above
(cffi:defcfun ("mfit_multi_create" %multi-create) :int (n-gpus :int) (dev-ids :pointer) (shard-mode :int) (out :pointer))
(cffi:defcfun ("mfit_multi_destroy" %multi-destroy) :void (m :pointer))
(cffi:defcfun ("mfit_multi_last_error" %multi-last-error) :string (m :pointer))
(cffi:defcfun ("mfit_multi_reserve" %multi-reserve) :int (m :pointer) (max-chains :int64))
(cffi:defcfun ("mfit_multi_problem_define" %multi-problem-define) :int (m :pointer) (n-datasets :int) (n-params :int))
(cffi:defcfun ("mfit_multi_dataset_upload" %multi-dataset-upload) :int
  (m :pointer) (d :int) (x :pointer) (x-dims :int) (y :pointer) (sigma :pointer) (n :int64))
(cffi:defcfun ("mfit_multi_model_builtin" %multi-model-builtin) :int
  (m :pointer) (d :int) (family :int) (degree :int) (param-index :pointer) (n-idx :int))
(cffi:defcfun ("mfit_multi_model_sexpr" %multi-model-sexpr) :int
  (m :pointer) (d :int) (lambda-text :string) (key-names :pointer) (n-keys :int))
(cffi:defcfun ("mfit_multi_likelihood_set" %multi-likelihood-set) :int (m :pointer) (d :int) (kind :int))
(cffi:defcfun ("mfit_multi_likelihood_sexpr" %multi-likelihood-sexpr) :int (m :pointer) (d :int) (lambda-text :string))
(cffi:defcfun ("mfit_multi_prior_flat" %multi-prior-flat) :int (m :pointer) (d :int))
(cffi:defcfun ("mfit_multi_prior_bounds" %multi-prior-bounds) :int
  (m :pointer) (d :int) (n :int) (idx :pointer) (lo :pointer) (hi :pointer))
(cffi:defcfun ("mfit_multi_prior_sexpr" %multi-prior-sexpr) :int
  (m :pointer) (d :int) (lambda-text :string) (key-names :pointer) (n-keys :int))
(cffi:defcfun ("mfit_multi_chains_init" %multi-chains-init) :int
  (m :pointer) (theta0 :pointer) (w :int64) (seed :uint64) (trace-capacity :int64))
(cffi:defcfun ("mfit_multi_run_begin" %multi-run-begin) :int (m :pointer) (cfg :pointer) (l-matrix :pointer))
(cffi:defcfun ("mfit_multi_run_advance" %multi-run-advance) :int
  (m :pointer) (max-steps :int64) (steps-done :pointer) (finished :pointer))
(cffi:defcfun ("mfit_multi_run_many_steps" %multi-run-many-steps) :int (m :pointer) (n :int64) (l-matrix :pointer))
(cffi:defcfun ("mfit_multi_request_stop" %multi-request-stop) :int (m :pointer))
(cffi:defcfun ("mfit_multi_trace_read" %multi-trace-read) :int
  (m :pointer) (take :int64) (theta :pointer) (prob :pointer) (n-valid :pointer))
(cffi:defcfun ("mfit_multi_stats_read" %multi-stats-read) :int (m :pointer) (stats :pointer))
(cffi:defcfun ("mfit_multi_best_read" %multi-best-read) :int (m :pointer) (theta :pointer) (prob :pointer))
(cffi:defcfun ("mfit_multi_last_read" %multi-last-read) :int (m :pointer) (theta :pointer) (prob :pointer))

(cffi:defcfun ("mfit_multi_trace_analytics" %multi-trace-analytics) :int
  (m :pointer) (take :int64) (most-likely-theta :pointer) (most-likely-prob :pointer) (median :pointer)
  (acceptance :pointer) (covariance :pointer) (l-matrix :pointer))
(cffi:defcfun ("mfit_multi_fit_envelope" %multi-fit-envelope) :int
  (m :pointer) (d :int) (take :int64) (which-solution :int) (x-fit :pointer) (n-fit :int) (y-fit :pointer)
  (y-max :pointer) (y-min :pointer) (params :pointer) (n-distinct :pointer))
(cffi:defcfun ("mfit_multi_walker_modify" %multi-walker-modify) :int (m :pointer) (mode :int) (number :int64))

(defvar *multi* nil "True while the handle in use is an mfit_multi (a walker made with :devices).")
(defmacro api (single multi handle &rest args)
  "Call SINGLE on a one-GPU context or its twin MULTI on a several-GPU handle."
  `(if *multi* (,multi ,handle ,@args) (,single ,handle ,@args)))
(defmacro with-walker-api ((walker) &body body)
  `(let ((*multi* (gpu-walker-multi ,walker))) ,@body))
(defun one-gpu-only (walker what)
  (when (gpu-walker-multi walker)
    (error "~a reads or edits the walk on one GPU; create the walker with :device for it" what)))

(defmacro with-traps-masked (&body body)
  #+sbcl `(sb-int:with-float-traps-masked (:invalid :divide-by-zero :overflow :inexact :underflow) ,@body)
  #-sbcl `(progn ,@body))

(define-condition mfit-error (error)
  ((code :initarg :code :reader mfit-error-code) (message :initarg :message :reader mfit-error-message))
  (:report (lambda (c s) (format s "mfit error ~d: ~a" (mfit-error-code c) (mfit-error-message c)))))

(defmacro chk (ctx form)
  "Signal a Lisp error for a negative mfit_status (the reference signals `error` in clean-data :819,:825)."
  (let ((rc (gensym)))
    `(let ((,rc (with-traps-masked ,form)))
       (unless (zerop ,rc) (error 'mfit-error :code ,rc :message (api %last-error %multi-last-error ,ctx)))
       ,rc)))

;;; ---------------------------------------------------------------------------------- plugin-slot descriptors
(defstruct gpu-fn text)                       ; printed lambda, lowered by the library (csrc/lower.cpp)
(defstruct gpu-family name degree keys)

(defun print-form (form)
  "The text handed to mfit_model_sexpr / mfit_prior_sexpr: symbols by name (no package prefixes, lower case,
keywords keep their colon), numbers in standard syntax (doubles keep their d0, ratios stay ratios)."
  (with-output-to-string (out)
    (labels ((emit (f)
               (cond ((null f) (write-string "nil" out))
                     ((consp f)
                      (write-char #\( out)
                      (loop for (item . rest) on f
                            do (emit item)
                               (cond ((consp rest) (write-char #\Space out))
                                     (rest (write-string " . " out) (emit rest))))
                      (write-char #\) out))
                     ((keywordp f) (write-char #\: out) (write-string (string-downcase (symbol-name f)) out))
                     ((symbolp f) (write-string (string-downcase (symbol-name f)) out))
                     (t (with-standard-io-syntax (prin1 f out))))))
      (emit form))))

(defmacro gpu-lambda (lambda-list &body body)
  "The lowering macro: captures a reference-style model or prior lambda as source text.  The library
translates it to CUDA C and NVRTC compiles it for sm_100a; it never runs in Lisp."
  `(make-gpu-fn :text ,(print-form `(lambda ,lambda-list ,@body))))

(defparameter *families*                       ; ids of include/mfit_cuda.h; theta order as documented there
  '((:poly . 0) (:expo . 1) (:sinus . 2) (:lorentz . 3) (:gauss-lin . 4) (:nv-multi . 5)))

(defun gpu-builtin (name keys &key (degree 0))
  (unless (assoc name *families*) (error "unknown family ~s" name))
  (make-gpu-family :name name :degree degree :keys keys))

;;; ---------------------------------------------------------------------------------- foreign staging helpers
(defmacro with-doubles ((ptr seq) &body body)
  "Copy a sequence of reals (or NIL -> null pointer) into a foreign double array for the duration of BODY
(to-double-floats :833-835 happens in the coerce)."
  (let ((s (gensym)) (n (gensym)) (i (gensym)))
    `(let* ((,s ,seq) (,n (length ,s)))
       (if (null ,s)
           (let ((,ptr (cffi:null-pointer))) ,@body)
           (cffi:with-foreign-object (,ptr :double (max 1 ,n))
             (let ((,i 0))
               (map nil (lambda (v) (setf (cffi:mem-aref ,ptr :double ,i) (coerce v 'double-float)) (incf ,i)) ,s))
             ,@body)))))

(defmacro with-key-names ((ptr keys) &body body)
  (let ((strs (gensym)) (i (gensym)))
    `(let ((,strs (mapcar (lambda (k) (cffi:foreign-string-alloc (string-downcase (symbol-name k)))) ,keys)))
       (unwind-protect
            (cffi:with-foreign-object (,ptr :pointer (max 1 (length ,strs)))
              (let ((,i 0)) (dolist (s ,strs) (setf (cffi:mem-aref ,ptr :pointer ,i) s) (incf ,i)))
              ,@body)
         (mapc #'cffi:foreign-string-free ,strs)))))

(defun doubles-from (ptr n)
  (let ((v (make-array n :element-type 'double-float)))
    (dotimes (i n v) (setf (aref v i) (cffi:mem-aref ptr :double i)))))

(defun lower-to-cuda (fn keys &key (kind 0))
  "The CUDA C a gpu-lambda lowers to (kind 0 = :function, 1 = :log-prior).  Needs no GPU."
  (with-key-names (kp keys)
    (cffi:with-foreign-objects ((err :char 1024))
      (let ((need (%lower-sexpr (gpu-fn-text fn) kp (length keys) kind (cffi:null-pointer) 0 err 1024)))
        (when (minusp need) (error 'mfit-error :code need :message (cffi:foreign-string-to-lisp err)))
        (cffi:with-foreign-object (out :char need)
          (%lower-sexpr (gpu-fn-text fn) kp (length keys) kind out need err 1024)
          (cffi:foreign-string-to-lisp out))))))

;;; ---------------------------------------------------------------------------------- walker
(defstruct gpu-walker ctx multi param-style single-key function param-keys n-walkers n-params data data-error log-liklihood log-prior)

(defvar mfit-walker-estop nil "Set to T (e.g. from another thread) to stop a run between launches, :860.")

(defun force-list (x) (if (listp x) x (list x)))          ; :755-759
(defun plist-keys (p) (loop for k in p by #'cddr collect k))
(defun plist-values (p) (loop for v in (cdr p) by #'cddr collect v))

;;; :single-item parameter style (:8-14, :1078-1088, :1153-1155): ONE key whose value is a list, a vector or an n x 1 array.
;;; The device sees it flattened -- element k is the parameter named NAME[k], which the lowering resolves from
;;; (elt NAME k) / (aref NAME k) / (aref NAME k 0) / (nth k NAME) in the :function text.
(defun single-item-p (params)
  (and (= (length params) 2) (second params) (typep (second params) '(and (or list vector array) (not string)))))
(defun sequence-values (v)
  (etypecase v
    (list v)
    (vector (coerce v 'list))
    (array (loop for i below (array-dimension v 0) collect (aref v i 0)))))
(defun flat-param-keys (params)
  (if (single-item-p params)
      (loop for k below (length (sequence-values (second params)))
            collect (intern (format nil "~a[~d]" (symbol-name (first params)) k) :keyword))
      (plist-keys params)))
(defun flat-param-values (params)
  (if (single-item-p params) (sequence-values (second params)) (plist-values params)))
(defun shape-params (walker plist)
  "A flat (key value ...) list as the device returns it -> the reference's shape: (:name (v0 v1 ...)) for :single-item."
  (if (eq (gpu-walker-param-style walker) :single-item)
      (list (gpu-walker-single-key walker) (plist-values plist))
      plist))

(defun clean-data (data n-functions)
  "clean-data :807-825: one (x y) set, or a list of them, one per function."
  (let ((sets (if (and (consp (first data)) (consp (first (first data)))) data (list data))))
    (unless (= (length sets) n-functions)
      (error "Number of data sets (~d) does not match number of functions (~d)" (length sets) n-functions))
    sets))

(defun clean-data-error (err data)
  "clean-data-error :774-805: a full per-point list per data set is kept; anything else is the FIRST number found."
  (labels ((first-number (e) (if (consp e) (first-number (first e)) e)))
    (let ((errs (if (and (consp err) (consp (first err)) (= (length err) (length data))) err nil)))
      (loop for d in data for k from 0
            for y = (second d)
            collect (cond ((and errs (= (length (nth k errs)) (length y))) (nth k errs))
                          ((and (consp err) (numberp (first err)) (= (length err) (length y))) err)
                          (t (make-list (length y) :initial-element (first-number err))))))))

(defun likelihood-kind (x)
  (ecase (or x :log-liklihood-normal)
    (:log-liklihood-normal 0) (:log-liklihood-normal-cutoff 1) (:log-poisson 2)))

(defun create-log-liklihood-function (log-liklihood-function)
  "create-log-liklihood-function :402-417 for the GPU path: LOG-LIKLIHOOD-FUNCTION is a per-point
(gpu-lambda (y model error) ...); the result goes into :log-liklihood.  The lambda is lowered to CUDA C and compiled into
the likelihood kernel together with the data set's :function; as in the reference (:416) ERROR is the data set's whole
stddev list -- read it with (elt error k), k a literal."
  (check-type log-liklihood-function gpu-fn)
  log-liklihood-function)

(defun set-likelihood (ctx d ll)
  (if (gpu-fn-p ll)
      (chk ctx (api %likelihood-sexpr %multi-likelihood-sexpr ctx d (gpu-fn-text ll)))
      (chk ctx (api %likelihood-set %multi-likelihood-set ctx d (likelihood-kind ll)))))

(defun set-function (ctx d f keys)
  (etypecase f
    (gpu-family
     (let ((idx (mapcar (lambda (k) (or (position k keys) (error "key ~s of function ~d is not in :params" k d)))
                        (gpu-family-keys f))))
       (cffi:with-foreign-object (ip :int32 (max 1 (length idx)))
         (loop for v in idx for i from 0 do (setf (cffi:mem-aref ip :int32 i) v))
         (chk ctx (api %model-builtin %multi-model-builtin ctx d (cdr (assoc (gpu-family-name f) *families*)) (gpu-family-degree f) ip (length idx))))))
    (gpu-fn (with-key-names (kp keys) (chk ctx (api %model-sexpr %multi-model-sexpr ctx d (gpu-fn-text f) kp (length keys)))))))

(defun set-prior (ctx d p keys)
  (cond ((or (null p) (eq p :log-prior-flat)) (chk ctx (api %prior-flat %multi-prior-flat ctx d)))
        ((gpu-fn-p p) (with-key-names (kp keys) (chk ctx (api %prior-sexpr %multi-prior-sexpr ctx d (gpu-fn-text p) kp (length keys)))))
        (t (let ((n (length p)))                   ; ((:key lo hi) ...), a missing key reads 0d0 like getf :353
             (cffi:with-foreign-object (ip :int32 (max 1 n))
               (loop for (k) in p for i from 0 do (setf (cffi:mem-aref ip :int32 i) (or (position k keys) -1)))
               (with-doubles (lo (mapcar #'second p))
                 (with-doubles (hi (mapcar #'third p))
                   (chk ctx (api %prior-bounds %multi-prior-bounds ctx d n ip lo hi)))))))))

(defun walker-create (&key function data params data-error log-liklihood log-prior param-bounds
                        (n-walkers 1) theta0 (seed 0) (device 0) devices (shard :chains) (trace-capacity 0))
  "walker-create :1132-1163.  GPU-side extras: N-WALKERS chains start at PARAMS, or THETA0 (a list of W value
lists in :params order) gives each chain its own start; SEED keys the Philox streams.  DEVICES (a list of GPU ordinals)
shards the walker set over several GPUs of this box inside the library: SHARD :chains splits the chains, :datasets puts
data set d of a global fit on GPU d mod n; results are bit-identical to one GPU."
  (declare (ignorable param-bounds))
  (let* ((*multi* (and devices t))
         (function (force-list function))
         (d-count (length function))
         (data (clean-data data d-count))
         (data-error (clean-data-error (if data-error data-error 1) data))
         (keys (flat-param-keys params))
         (p (length keys))
         (ll (if (consp log-liklihood) log-liklihood (make-list d-count :initial-element log-liklihood)))
         (lp (if (and (consp log-prior) (= (length log-prior) d-count) (not (keywordp (first (first log-prior)))))
                 log-prior (make-list d-count :initial-element log-prior)))
         (theta0 (or theta0 (make-list n-walkers :initial-element (flat-param-values params))))
         (w (length theta0))
         ctx)
    (cffi:with-foreign-objects ((out :pointer) (ids :int (max 1 (length devices))))
      (loop for g in devices for i from 0 do (setf (cffi:mem-aref ids :int i) g))
      (let ((rc (with-traps-masked
                  (if *multi*
                      (%multi-create (length devices) ids (ecase shard (:chains 0) (:datasets 1)) out)
                      (%ctx-create device (cffi:null-pointer) out)))))
        (unless (zerop rc)
          (error 'mfit-error :code rc :message (api %last-error %multi-last-error (cffi:null-pointer)))))
      (setf ctx (cffi:mem-ref out :pointer)))
    (when (and *multi* (eq shard :datasets)) (chk ctx (%multi-reserve ctx (max w 4096))))
    (chk ctx (api %problem-define %multi-problem-define ctx d-count p))
    (loop for d from 0 for f in function for set in data for e in data-error do
      (let* ((xs (first set))
             ;; multi-dimensional inputs (:1137): every x is itself a sequence -> component-major for the device
             (x-dims (if (and (plusp (length xs)) (typep (elt xs 0) 'sequence)) (length (elt xs 0)) 0))
             (flat (if (plusp x-dims)
                       (loop for k below x-dims append (map 'list (lambda (xi) (elt xi k)) xs))
                       xs)))
        (with-doubles (x flat)
          (with-doubles (y (second set))
            (with-doubles (s e)
              (chk ctx (cond (*multi* (%multi-dataset-upload ctx d x x-dims y s (length (second set))))
                             ((plusp x-dims) (%dataset-upload-xd ctx d x x-dims y s (length (second set))))
                             (t (%dataset-upload ctx d x y s (length (second set))))))))))
      (set-function ctx d f keys)
      (set-likelihood ctx d (nth d ll))
      (set-prior ctx d (nth d lp) keys))
    (with-doubles (t0 (apply #'append theta0))               ; W x P row-major
      (chk ctx (if *multi*
                   (%multi-chains-init ctx t0 w seed trace-capacity)
                   (%chains-init ctx t0 w seed 0 trace-capacity))))
    (make-gpu-walker :ctx ctx :multi *multi* :param-style (if (single-item-p params) :single-item :multiple-kwargs)
                     :single-key (first params) :function function :param-keys keys :n-walkers w :n-params p :data data
                     :data-error data-error :log-liklihood ll :log-prior lp)))

(defun walker-set-create (&key function x ys params data-error log-liklihood log-prior (seed 0) (device 0) (trace-capacity 0))
  "A walker set with per-walker data -- what file->nv-walkers builds (nv-specific.lisp:50-66): one walker per y list in
YS, all over the common X, same :function / :log-liklihood / :log-prior, every walker starting at PARAMS.  DATA-ERROR is
a number or one list as long as X.  The walkers advance together on the GPU, one warp each."
  (let* ((keys (plist-keys params)) (p (length keys)) (w (length ys)) (n (length x))
         (sigma (if (listp data-error) data-error (make-list n :initial-element (or data-error 1))))
         ctx)
    (cffi:with-foreign-object (out :pointer)
      (let ((rc (with-traps-masked (%ctx-create device (cffi:null-pointer) out))))
        (unless (zerop rc) (error 'mfit-error :code rc :message (%last-error (cffi:null-pointer)))))
      (setf ctx (cffi:mem-ref out :pointer)))
    (chk ctx (%problem-define ctx 1 p))
    (with-doubles (xp x)
      (with-doubles (yp (apply #'append ys))                 ; W x n row-major
        (with-doubles (sp sigma)
          (chk ctx (%dataset-upload-per-walker ctx 0 xp yp sp n w 1 1)))))
    (set-function ctx 0 function keys)
    (chk ctx (%likelihood-set ctx 0 (likelihood-kind log-liklihood)))
    (set-prior ctx 0 log-prior keys)
    (with-doubles (t0 (loop repeat w append (plist-values params)))
      (chk ctx (%chains-init ctx t0 w seed 0 trace-capacity)))
    (make-gpu-walker :ctx ctx :function (list function) :param-keys keys :n-walkers w :n-params p
                     :data (list x ys) :data-error sigma :log-liklihood (list log-liklihood) :log-prior (list log-prior))))

(defun walker-free (walker)
  (when (gpu-walker-ctx walker)
    (if (gpu-walker-multi walker) (%multi-destroy (gpu-walker-ctx walker)) (%ctx-destroy (gpu-walker-ctx walker)))
    (setf (gpu-walker-ctx walker) nil)))

(defun walker-adaptive-steps-full (walker &key (n 100000) (temperature 1d3) (auto :prob-settle)
                                            (sampling-optimization :covariance) max-walker-length l-matrix
                                            (poll-every 512))
  "walker-adaptive-steps-full :862-942, all chains advanced together on the GPU.  The run is issued in slices of
POLL-EVERY steps so that mfit-walker-estop (:860, :904) is honoured.  :auto :slope-settle is not supported."
  (setf mfit-walker-estop nil)
  (let ((ctx (gpu-walker-ctx walker)) (p (gpu-walker-n-params walker)) (*multi* (gpu-walker-multi walker)))
    (cffi:with-foreign-objects ((cfg '(:struct adapt-cfg)) (done :int64) (fin :int))
      (cffi:with-foreign-slots ((n temperature auto-mode sampling-optimization max-walker-length greedy reserved)
                                cfg (:struct adapt-cfg))
        (setf n (floor n) temperature (coerce temperature 'double-float)
              auto-mode (ecase auto ((nil) 0) (:prob-settle 1))
              sampling-optimization (ecase sampling-optimization (:covariance 0) (:best-value 1))
              max-walker-length (if max-walker-length (floor max-walker-length) 0) greedy 0 reserved 0))
      (with-doubles (lm (when l-matrix (loop for i below p append (loop for j below p collect (aref l-matrix i j)))))
        (chk ctx (if *multi*
                     (%multi-run-begin ctx cfg lm)
                     (%run-begin ctx cfg lm (cffi:null-pointer) (cffi:null-pointer) 0))))
      (loop
        (when mfit-walker-estop (api %request-stop %multi-request-stop ctx))
        (chk ctx (api %run-advance %multi-run-advance ctx poll-every done fin))
        (unless (zerop (cffi:mem-ref fin :int)) (return)))))
  walker)

(defun walker-adaptive-steps (walker &optional (n 30000))                               ; :946-947
  (walker-adaptive-steps-full walker :n n :temperature 10 :auto :prob-settle))

(defun walker-many-steps (walker n &optional l-matrix)                                  ; :849-853
  "walker-many-steps :849-853.  Without L-MATRIX every chain gets the reference's default (:851),
(diagonal-covariance (scale-plist 1e-2 median-params)), from its own walk (medians computed on the device)."
  (let ((ctx (gpu-walker-ctx walker)) (p (gpu-walker-n-params walker)) (w (gpu-walker-n-walkers walker))
        (*multi* (gpu-walker-multi walker)))
    (if l-matrix
        (with-doubles (lm (loop for i below p append (loop for j below p collect (aref l-matrix i j))))
          (chk ctx (if *multi*
                       (%multi-run-many-steps ctx n lm)
                       (%run-many-steps ctx n lm (cffi:null-pointer) (cffi:null-pointer) 0))))
        (progn
          (cffi:with-foreign-objects ((med :double (* w p)) (lms :double (* w p p)))
            (let ((null (cffi:null-pointer)))
              (chk ctx (api %trace-analytics %multi-trace-analytics ctx 1000000000 null null med null null null))
              (dotimes (c w)
                (dotimes (i p)
                  (dotimes (j p)
                    (setf (cffi:mem-aref lms :double (+ (* c p p) (* i p) j))
                          (if (= i j) (* (coerce 1e-2 'double-float) (cffi:mem-aref med :double (+ (* c p) i))) 0d0)))))
              (chk ctx (%run-many-steps-per-chain ctx n lms null null 0)))))))
  walker)

(defun walker-sample-region (walker &optional (initial-scale 1d-3))                     ; :949-969
  (one-gpu-only walker "walker-sample-region")
  (chk (gpu-walker-ctx walker)
       (%run-sample-region (gpu-walker-ctx walker) (coerce initial-scale 'double-float) (cffi:null-pointer) 0))
  walker)

(defun walker-force-take-step (walker)                                                  ; :1124-1129
  (one-gpu-only walker "walker-force-take-step")
  (chk (gpu-walker-ctx walker) (%force-take-step (gpu-walker-ctx walker)))
  walker)

(defun read-step (walker reader chain)
  (let* ((ctx (gpu-walker-ctx walker)) (w (gpu-walker-n-walkers walker)) (p (gpu-walker-n-params walker))
         (*multi* (gpu-walker-multi walker)))
    (cffi:with-foreign-objects ((th :double (* w p)) (pr :double w))
      (chk ctx (funcall reader ctx th pr))
      (values (shape-params walker (loop for k in (gpu-walker-param-keys walker) for i from 0
                                         append (list k (cffi:mem-aref th :double (+ (* chain p) i)))))
              (cffi:mem-aref pr :double chain)))))

(defun walker-get (walker &key (get :steps) take param (chain 0))
  "walker-get :487-543 for the modes that read the trace of one CHAIN, newest first."
  (let* ((ctx (gpu-walker-ctx walker)) (w (gpu-walker-n-walkers walker)) (p (gpu-walker-n-params walker))
         (keys (gpu-walker-param-keys walker)) (*multi* (gpu-walker-multi walker))
         (best-reader (if *multi* #'%multi-best-read #'%best-read))
         (last-reader (if *multi* #'%multi-last-read #'%last-read)))
    (ecase get
      (:most-likely-params (values (read-step walker best-reader chain)))
      (:most-likely-step (read-step walker best-reader chain))
      (:last-step (read-step walker last-reader chain))
      ((:median-params :stddev-params :covariance-matrix :l-matrix)      ; :516-543, computed on the device for all chains
       (let ((take (or take 1000000000)) (null (cffi:null-pointer)))
         (cffi:with-foreign-objects ((med :double (* w p)) (mat :double (* w p p)))
           (ecase get
             (:median-params (chk ctx (api %trace-analytics %multi-trace-analytics ctx take null null med null null null)))
             (:covariance-matrix (chk ctx (api %trace-analytics %multi-trace-analytics ctx take null null null null mat null)))
             ((:l-matrix :stddev-params) (chk ctx (api %trace-analytics %multi-trace-analytics ctx take null null null null null mat))))
           (flet ((mat-at (i j) (cffi:mem-aref mat :double (+ (* chain p p) (* i p) j))))
             (ecase get
               (:median-params (loop for k in keys for i from 0
                                     append (list k (cffi:mem-aref med :double (+ (* chain p) i)))))
               (:stddev-params (loop for k in keys for i from 0 append (list k (mat-at i i))))
               ((:covariance-matrix :l-matrix)
                (let ((a (make-array (list p p) :element-type 'double-float)))
                  (dotimes (i p a) (dotimes (j p) (setf (aref a i j) (mat-at i j)))))))))))
      ((:steps :params :param :log-liklihoods :acceptance)
       (let ((take (or take (cffi:with-foreign-object (st '(:struct chain-stats) w)
                              (chk ctx (api %stats-read %multi-stats-read ctx st))
                              (cffi:foreign-slot-value (cffi:mem-aptr st '(:struct chain-stats) chain)
                                                       '(:struct chain-stats) 'length)))))
         (cffi:with-foreign-objects ((th :double (* w take p)) (pr :double (* w take)) (nv :int64 w))
           (chk ctx (api %trace-read %multi-trace-read ctx take th pr nv))
           (let* ((m (cffi:mem-aref nv :int64 chain))
                  (probs (loop for r below m collect (cffi:mem-aref pr :double (+ (* chain take) r))))
                  (rows (loop for r below m
                              collect (loop for k in keys for i from 0
                                            append (list k (cffi:mem-aref th :double (+ (* (+ (* chain take) r) p) i)))))))
             (ecase get
               (:steps (mapcar #'cons probs (mapcar (lambda (row) (shape-params walker row)) rows)))
               (:params (mapcar (lambda (row) (shape-params walker row)) rows))
               (:param (mapcar (lambda (row) (getf row param)) rows))
               (:log-liklihoods probs)
               (:acceptance (/ (1+ (loop for (a b) on probs while b count (/= a b))) m))))))))))   ; :506-508

(defun walker-get-data-and-fit (walker &key (take 1000) (x-column 0) (y-column 1) (fn-number 0)
                                          (which-solution :most-likely) x-shift y-shift (chain 0))
  "walker-get-data-and-fit :1231-1255 for one CHAIN -> (x-fit max-ys min-ys y-fit x-data y-data params).  The fit
curve and the envelope over the best 66 % of the walk are evaluated on the device for every chain at once."
  (let* ((ctx (gpu-walker-ctx walker)) (w (gpu-walker-n-walkers walker)) (p (gpu-walker-n-params walker))
         (*multi* (gpu-walker-multi walker))
         (data (elt (gpu-walker-data walker) fn-number))
         (x-data (elt data x-column)) (y-data (elt data y-column))
         (x-fit (let* ((lo (reduce #'min x-data)) (hi (reduce #'max x-data))           ; linspace :235-247, :len 1000
                       (step (/ (rational (- hi lo)) 999)))
                  (loop for n below 1000 collect (coerce (+ (rational lo) (* n step)) 'double-float))))
         (n-fit (length x-fit))
         (shift (lambda (s seq) (if s (map 'list (lambda (v) (+ s v)) seq) (coerce seq 'list)))))
    (cffi:with-foreign-objects ((xf :double n-fit) (yf :double (* w n-fit)) (ymax :double (* w n-fit))
                                (ymin :double (* w n-fit)) (par :double (* w p)))
      (loop for x in x-fit for i from 0 do (setf (cffi:mem-aref xf :double i) (coerce x 'double-float)))
      (with-traps-masked
        (chk ctx (api %fit-envelope %multi-fit-envelope ctx fn-number take (if (eq which-solution :most-likely) 0 1) xf n-fit
                      yf ymax ymin par (cffi:null-pointer))))
      (flet ((row (ptr) (loop for i below n-fit collect (cffi:mem-aref ptr :double (+ (* chain n-fit) i)))))
        (list (funcall shift x-shift x-fit)
              (funcall shift y-shift (row ymax))
              (funcall shift y-shift (row ymin))
              (funcall shift y-shift (row yf))
              (funcall shift x-shift x-data)
              (funcall shift y-shift y-data)
              (loop for k in (gpu-walker-param-keys walker) for i from 0
                    append (list k (cffi:mem-aref par :double (+ (* chain p) i)))))))))

(defun walker-save (walker filename &key take (chain 0))
  "walker-save :971-1001 revived: one readable s-expression per file."
  (with-open-file (out filename :direction :output :if-exists :supersede)
    (with-standard-io-syntax
      (print (list :fn (mapcar (lambda (f) (if (gpu-fn-p f) (gpu-fn-text f) (gpu-family-name f))) (gpu-walker-function walker))
                   :param-keys (gpu-walker-param-keys walker)
                   :data (gpu-walker-data walker) :stddev (gpu-walker-data-error walker)
                   :log-liklihood (gpu-walker-log-liklihood walker)
                   :walker (walker-get walker :get :steps :take take :chain chain))
             out)))
  filename)

(defun walker-load (filename &key function log-liklihood log-prior quiet (seed 0) (device 0))
  "walker-load :989-1001 revived for a file written by walker-save: prints the saved names as *Recommendations*; with
:function given, rebuilds the walker (walker-create on the newest step, :999) and restores its walk on the device
(:add-walks :556-563 as the loader means it: the walk, the last step and the most likely step)."
  (with-open-file (in filename :direction :input)
    (with-standard-io-syntax
      (let* ((full (read in))
             (data (getf full :data))
             (data-error (getf full :stddev))
             (walks (getf full :walker))                      ; ((prob . plist) ...), newest first
             (keys (getf full :param-keys))
             (dummy-params (cdr (first walks))))
        (unless quiet
          (format t "*Recommendations*~%function: ~s~%log-liklihood: ~s~%" (getf full :fn) (getf full :log-liklihood)))
        (when function
          (let* ((walker (walker-create :function function :data data :params dummy-params :data-error data-error
                                        :log-liklihood log-liklihood :log-prior log-prior :seed seed :device device))
                 (ctx (gpu-walker-ctx walker)) (rows (length walks)) (p (length keys)))
            (cffi:with-foreign-objects ((th :double (* rows p)) (pr :double rows) (nv :int64 1))
              (loop for (prob . plist) in walks for r from 0 do
                (setf (cffi:mem-aref pr :double r) (coerce prob 'double-float))
                (loop for k in keys for i from 0 do
                  (setf (cffi:mem-aref th :double (+ (* r p) i)) (coerce (getf plist k) 'double-float))))
              (setf (cffi:mem-aref nv :int64 0) rows)
              (with-traps-masked
                (chk ctx (%chains-restore ctx rows th pr nv (cffi:null-pointer) (cffi:null-pointer) (cffi:null-pointer)))))
            walker))))))

(defun walker-set-get (walker &key (get :steps) take param)
  "walker-set-get :1029-1030: a gpu-walker is a set of chains; one walker-get result per chain."
  (loop for c below (gpu-walker-n-walkers walker)
        collect (walker-get walker :get get :take take :param param :chain c)))

(defun walker-modify (walker &key (modify :keep-walks) burn-number keep-number)
  "walker-modify :546-580 on the device, every chain: :burn-walks :keep-walks :reset :reset-to-most-likely."
  (let ((ctx (gpu-walker-ctx walker)) (*multi* (gpu-walker-multi walker)))
    (chk ctx (ecase modify
               (:burn-walks (api %walker-modify %multi-walker-modify ctx 1 burn-number))
               (:keep-walks (api %walker-modify %multi-walker-modify ctx 2 keep-number))
               (:reset (api %walker-modify %multi-walker-modify ctx 3 0))
               (:reset-to-most-likely (api %walker-modify %multi-walker-modify ctx 4 0))))
    walker))

(defun mcmc-fit (&rest args &key function data params data-error log-liklihood log-prior param-bounds &allow-other-keys)
  "mcmc-fit :1165-1175"
  (declare (ignorable function data params data-error log-liklihood log-prior param-bounds))
  (walker-adaptive-steps (apply #'walker-create args)))
